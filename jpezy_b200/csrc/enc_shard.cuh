// enc_shard.cuh -- one image encoded by several GPUs, sharded by MCU rows (BASELINE.json config 5).
//
// The reference has no counterpart (it is single threaded); what must be preserved is its OUTPUT: one restart-less
// entropy-coded segment (src/encoder/jpezy_encoder.hpp:58-67).  Two image-wide prefix dependencies cross the shards:
// the DC predictors pre_DC[3] (:180-181) and the bit cursor of the bit writer.  Per rank k:
//   A  transform the local MCU rows; publish the last quantised DCs (Y3, Cb, Cr of the last MCU)        -> all-gather #1
//   B  code lengths with the previous rank's DCs as predictors, local scan, scatter at LOCAL bit 0;
//      publish {T_k = local bit count, head_k = first 8 local bits}                                       -> all-gather #2
//   C  B_k = sum_{j<k} T_j.  Rank k owns the globally aligned bytes [ceil(B_k/8), ceil(B_{k+1}/8)): the byte it shares
//      with rank k+1 is completed with head_{k+1} (the last rank: with the pad bits of write_eoi).  Owned byte i is
//      the 8 local bits starting at 8i + d, d = 8*ceil(B_k/8) - B_k.  Count the 0xFF bytes among them (stuffing depends
//      on the GLOBAL byte alignment, so it can only happen now); publish owned + stuffed byte count          -> all-gather #3
//   D  byte base = sum of the previous ranks' counts; write the stuffed bytes straight into the stitched output, which
//      may be a peer-mapped buffer on another GPU (stores travel over NVLink; no staging copy).
// Every shard holds at least one MCU (>= 24 bits), so a byte is shared by at most two ranks.
#pragma once
#include "enc_entropy.cuh"

namespace jz {

struct ShardInfo {       // all-gather #2 record
    uint64_t bits;       // T_k
    uint64_t head;       // first 8 bits of the local stream (local byte 0)
};

struct ShardGeom {       // per rank, device resident (written by k_shard_geom)
    uint64_t bit_base;   // B_k
    uint64_t first_own;  // ceil(B_k / 8): global index of the first owned byte
    uint64_t nown;       // owned bytes (before stuffing)
    uint32_t d;          // local bit offset of owned byte 0 (0..7)
    uint32_t fits;       // local stream fits its scratch buffer
};

struct ShardParams {
    EntParams e;                 // local entropy state (nimg == 1)
    const ShardInfo* all_info;   // [nranks]
    const uint64_t* all_bytes;   // [nranks] owned + stuffed bytes per rank
    uint32_t rank, nranks;
    int pad_ones;                // fill bits of the very last byte (JPEZYB200_OPT_PAD_ONES)
    ShardGeom* geom;
    uint64_t* out_bytes;         // this rank's owned + stuffed byte count (all-gather #3 input)
    uint8_t* stage;              // local staging of this rank's stuffed bytes, same 16-byte phase as their place in dst
    size_t stage_cap;
    uint8_t* dst;                // stitched output (possibly on a peer GPU)
    size_t dst_cap;
    uint64_t* total_bytes;       // optional: sum over all ranks (written by every rank, same value)
    int32_t* overflow;           // set to 1 when the stitched stream does not fit dst_cap / local scratch
};

// last quantised DCs of the shard = predictors of the next shard's first Y / Cb / Cr blocks
__global__ void k_shard_last_dc(const int16_t* __restrict__ coefs, uint32_t nmcu, int32_t* __restrict__ last_dc)
{
    pdl_wait();
    if (threadIdx.x < 3) {
        const int blk = threadIdx.x == 0 ? 3 : 3 + threadIdx.x;   // Y3, Cb, Cr
        last_dc[threadIdx.x] = coefs[(size_t(nmcu) - 1) * 384 + blk * 64];
    }
}

__global__ void k_shard_info(const EntParams p, ShardInfo* __restrict__ info)
{
    pdl_wait();
    if (threadIdx.x == 0) {
        const uint64_t bits = p.img_bits[0];
        info->bits = bits;
        info->head = ((bits + 7) / 8 + 4 <= p.uslot) ? uint64_t(p.ustream[0]) : 0ull;
    }
}

__global__ void k_shard_geom(const ShardParams p)
{
    pdl_wait();
    if (threadIdx.x != 0) return;
    uint64_t base = 0;
    for (uint32_t j = 0; j < p.rank; ++j) base += p.all_info[j].bits;
    const uint64_t T = p.all_info[p.rank].bits;
    const bool last = p.rank + 1 == p.nranks;
    ShardGeom g;
    g.bit_base = base;
    g.first_own = (base + 7) / 8;
    g.nown = (base + T + 7) / 8 - g.first_own;
    g.d = uint32_t(g.first_own * 8 - base);
    g.fits = (T + 7) / 8 + 20 <= p.e.uslot;
    *p.geom = g;
    if (!g.fits) {
        *p.overflow = 1;
        return;
    }
    // complete the shared byte: the next rank's leading bits, or the pad bits of write_eoi (src/encoder/jpezy_writer.hpp:101-105)
    const uint32_t head = last ? (p.pad_ones ? 0xffu : 0u) : uint32_t(p.all_info[p.rank + 1].head & 0xffu);
    const uint32_t sh = uint32_t(T & 7u);
    uint8_t* L = p.e.ustream + T / 8;
    L[0] |= uint8_t(head >> sh);
    if (sh) L[1] |= uint8_t(head << (8 - sh));
}

// 16 owned bytes starting at owned index i0 (multiple of 16): local bits [8*i0 + d, ...)
__device__ __forceinline__ void load_owned16(const uint8_t* __restrict__ L, uint64_t i0, uint32_t d, uint8_t (&out)[16])
{
    const uint4 a = *reinterpret_cast<const uint4*>(L + i0);
    const uint32_t nxt = L[i0 + 16];
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
    uint32_t cur = w[0] & 0xffu;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t nb = i < 15 ? (w[(i + 1) >> 2] >> (((i + 1) & 3) * 8)) & 0xffu : nxt;
        out[i] = uint8_t(((cur << d) | (nb >> (8 - d))) & 0xffu);   // d == 0: nb >> 8 == 0
        cur = nb;
    }
}

__global__ void __launch_bounds__(kStuffThreads) k_shard_ff_count(const ShardParams p)
{
    pdl_wait();
    __shared__ uint32_t s_warp[kStuffThreads / 32];
    const ShardGeom g = *p.geom;
    if (!g.fits) return;
    const uint32_t nch = uint32_t((g.nown + kStuffChunk - 1) / kStuffChunk);
    for (uint32_t ch = blockIdx.x; ch < nch; ch += gridDim.x) {
        const uint64_t i0 = (uint64_t(ch) * kStuffThreads + threadIdx.x) * 16;
        uint32_t c = 0;
        if (i0 < g.nown) {
            uint8_t b[16];
            load_owned16(p.e.ustream, i0, g.d, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) c += (b[i] == 0xffu && i0 + i < g.nown) ? 1u : 0u;
        }
        uint32_t total;
        block_scan_excl(c, s_warp, &total);
        if (threadIdx.x == 0) p.e.ff_sum[ch] = total;
    }
}

__global__ void __launch_bounds__(1024) k_shard_scan_ff(const ShardParams p)
{
    pdl_wait();
    __shared__ uint32_t s_warp[32];
    __shared__ uint64_t s_carry;
    const ShardGeom g = *p.geom;
    if (!g.fits) {
        // this rank's stream did not fit its scratch: UINT64_MAX travels through all-gather #3, so that EVERY rank raises its
        // overflow flag in k_shard_stuff_write (a silently short segment on rank 0 would be a corrupt file)
        if (threadIdx.x == 0) *p.out_bytes = ~0ull;
        return;
    }
    const uint32_t nch = uint32_t((g.nown + kStuffChunk - 1) / kStuffChunk);
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < nch; c0 += 1024) {
        const uint32_t c = c0 + threadIdx.x;
        const uint32_t v = c < nch ? p.e.ff_sum[c] : 0u;
        uint32_t total;
        const uint32_t off = block_scan_excl(v, s_warp, &total);
        const uint64_t carry = s_carry;
        if (c < nch) p.e.ff_base[c] = carry + off;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *p.out_bytes = g.nown + s_carry;
}

__global__ void __launch_bounds__(kStuffThreads) k_shard_stuff_write(const ShardParams p)
{
    pdl_wait();
    __shared__ uint32_t s_warp[kStuffThreads / 32];
    const ShardGeom g = *p.geom;
    uint64_t byte_base = 0, total = 0;
    bool peer_overflow = false;
    for (uint32_t j = 0; j < p.nranks; ++j) {
        peer_overflow |= p.all_bytes[j] == ~0ull;          // some rank (maybe this one) ran out of scratch
        if (j < p.rank) byte_base += p.all_bytes[j];
        total += p.all_bytes[j];
    }
    const uint64_t mine = p.all_bytes[p.rank];
    const bool ok = !peer_overflow && g.fits && total <= p.dst_cap && mine + 32 <= p.stage_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (p.total_bytes) *p.total_bytes = total;
        if (!ok) *p.overflow = 1;
    }
    if (!ok) return;
    const uint32_t nch = uint32_t((g.nown + kStuffChunk - 1) / kStuffChunk);
    // the stuffed bytes are first written locally (byte stores), at the same position modulo 16 as in the stitched
    // stream, so that k_shard_push can move them to the (possibly remote) destination with aligned 16-byte stores
    uint8_t* dst = p.stage + (byte_base & 15u);
    for (uint32_t ch = blockIdx.x; ch < nch; ch += gridDim.x) {
        const uint64_t i0 = (uint64_t(ch) * kStuffThreads + threadIdx.x) * 16;
        uint8_t b[16];
        uint32_t c = 0;
        const bool live = i0 < g.nown;
        if (live) {
            load_owned16(p.e.ustream, i0, g.d, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) c += (b[i] == 0xffu && i0 + i < g.nown) ? 1u : 0u;
        }
        uint32_t tot;
        const uint32_t off = block_scan_excl(c, s_warp, &tot);
        if (!live) continue;
        uint64_t o = i0 + p.e.ff_base[ch] + off;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i0 + i >= g.nown) break;
            dst[o++] = b[i];
            if (b[i] == 0xffu) dst[o++] = 0;
        }
    }
}

// local staging -> this rank's place in the stitched stream: aligned 16-byte loads and stores (peer stores over NVLink
// when dst lives on another GPU), single bytes only at the two ragged ends
__global__ void __launch_bounds__(256) k_shard_push(const ShardParams p)
{
    pdl_wait();
    if (*p.overflow) return;
    uint64_t byte_base = 0;
    for (uint32_t j = 0; j < p.rank; ++j) byte_base += p.all_bytes[j];
    const uint64_t n = p.all_bytes[p.rank];
    const uint32_t ph = uint32_t(byte_base & 15u);
    const uint8_t* src = p.stage;                 // byte i of this rank's segment is at src[ph + i]
    uint8_t* dst = p.dst + (byte_base - ph);      // 16-byte aligned (dst itself is): byte i goes to dst[ph + i]
    const uint64_t lo = ph, hi = ph + n;          // [lo, hi) in staging coordinates
    const uint64_t a0 = (lo + 15) & ~uint64_t(15), a1 = hi & ~uint64_t(15);
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x, nth = uint64_t(gridDim.x) * blockDim.x;
    if (a0 < a1) {
        for (uint64_t c = a0 / 16 + tid; c < a1 / 16; c += nth) reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
        for (uint64_t i = lo + tid; i < a0; i += nth) dst[i] = src[i];
        for (uint64_t i = a1 + tid; i < hi; i += nth) dst[i] = src[i];
    } else {
        for (uint64_t i = lo + tid; i < hi; i += nth) dst[i] = src[i];
    }
}

}  // namespace jz
