// dec_entropy.cuh -- stage D1: byte-stuffed Huffman segment -> quantised coefficients.
//
// Replaces decode_huffman / decode_huffman_impl (src/decoder/jpezy_decoder.hpp:583-642) and the bit
// reader behind them (srook::io::jpeg::bifstream: MSB-first, the 0x00 after 0xFF dropped).  The
// reference decodes one bit at a time, strictly sequentially; jpezy's encoder writes no restart
// markers, so the whole image is ONE entropy-coded segment.  It is decoded in parallel with the
// self-synchronising scheme of Weissenberger & Schmidt: the un-stuffed stream is cut into
// subsequences of `sub_bits` bits, every thread decodes its subsequence from a guessed decoder
// state, then repeatedly re-decodes from its predecessor's end state until all end states are a
// fixed point (Huffman codes re-synchronise after a few code words).  A prefix sum of the
// per-subsequence block counts gives each subsequence its first block index; a last pass decodes
// once more and writes the coefficients; DC differences are turned into absolute values by a
// per-component prefix sum.
//
//   D0  k_unstuff_count / k_scan_u32 / k_unstuff_write     drop stuffed zeros (compaction scan)
//   D1a k_sync_decode (round 0, then rounds until no end state changes)
//   D1b k_scan_u32 over block counts
//   D1c k_write_coefs
//   D1d k_dc_sum / k_scan_dc / k_dc_apply                  DC prediction = prefix sum per component
#pragma once
#include "common.cuh"

namespace jz {

constexpr int kDecThreads = 256;
constexpr int kLutBits = 10;

// compact canonical decoder table for one DHT table, built on the host
struct HuffDecTab {
    uint16_t fast[1 << kLutBits];  // peek kLutBits bits -> (len << 8) | symbol, 0 = longer than kLutBits / invalid
    int32_t maxcode[18];           // maxcode[len] (left-aligned compare uses plain codes), -1 = none
    int32_t valptr[17];            // index into vals of the first code of this length minus its code
    uint8_t vals[256];
};

struct DecParams {
    // geometry
    uint32_t nblk;            // blocks per image
    uint32_t nmcu;
    // stuffed input
    const uint8_t* scan;      // [nimg][slot]
    size_t slot;
    const uint64_t* scan_bytes;   // device copy of the host array [nimg]
    // un-stuffed stream
    uint8_t* ustream;         // [nimg][uslot]  (uslot multiple of 16, 32 bytes of zero slack)
    size_t uslot;
    uint32_t nchunk;          // 4 KiB chunks per image (capacity)
    uint32_t* chunk_cnt;      // [nimg][nchunk] kept bytes per chunk
    uint64_t* chunk_base;     // [nimg][nchunk]
    uint64_t* ubytes;         // [nimg] un-stuffed byte count
    // synchronisation
    uint32_t sub_bits;        // subsequence length in bits (power of two >= 128)
    uint32_t nsub;            // subsequences per image (capacity)
    uint32_t* state_a;        // [nimg][nsub] packed end states, double buffered
    uint32_t* state_b;
    uint8_t* dirty_a;         // [nimg][nsub]
    uint8_t* dirty_b;
    uint32_t* sub_blk;        // [nimg][nsub] exclusive prefix of block counts
    unsigned long long* changed;   // device flag: number of end states that changed this round
    // tables: [0]=DC sel by comp class 0, [1]=DC class 1, [2]=AC class 0, [3]=AC class 1
    const HuffDecTab* tabs;
    // output
    int16_t* coefs;           // [nimg][nblk*64]
    size_t coef_stride;
    int32_t* status;          // [nimg] or nullptr
    // DC scan scratch
    int32_t* dc_part;         // [nimg][ndc_tiles][3]
    uint32_t ndc_tiles;
};

// packed end state: overshoot(6) | b(3) | z(6) | nblocks(12) | valid(1)
__device__ __forceinline__ uint32_t pack_state(uint32_t over, uint32_t b, uint32_t z, uint32_t n)
{
    return (over & 63u) | (b << 6) | (z << 9) | (min(n, 4095u) << 15);
}
constexpr uint32_t kStateSyncMask = (1u << 15) - 1u;   // (overshoot, b, z)

// ---- bit reader over the un-stuffed stream (big-endian bit order) -------------------------------
struct BitPeek {
    const uint8_t* base;
    __device__ __forceinline__ uint32_t peek32(uint64_t p) const
    {
        // 32 bits starting at bit p (stream has >= 8 bytes of slack after the data)
        const uint64_t byte = p >> 3;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (byte & ~uint64_t(3)));
        const uint32_t a = __byte_perm(__ldg(w), 0, 0x0123), b = __byte_perm(__ldg(w + 1), 0, 0x0123);
        const uint32_t sh = uint32_t(p & 31u);
        return __funnelshift_l(b, a, sh);
    }
};

__device__ __forceinline__ uint32_t huff_lookup(const HuffDecTab* __restrict__ t, const uint16_t* __restrict__ s_fast, uint32_t bits32)
{
    // returns (len << 8) | symbol, len = 0 when no code matches
    const uint32_t e = s_fast[bits32 >> (32 - kLutBits)];
    if (e) return e;
#pragma unroll 1
    for (int len = kLutBits + 1; len <= 16; ++len) {
        const int32_t code = int32_t(bits32 >> (32 - len));
        if (code <= t->maxcode[len]) return (uint32_t(len) << 8) | t->vals[(code + t->valptr[len]) & 255];
    }
    return 0;
}

// Decode from (p, b, z) until p >= end or the image's blocks are exhausted.  Sink receives
// (block_ordinal_within_this_call, zz_index, value) for every coefficient with a value field.
template <bool kWrite>
__device__ __forceinline__ void decode_span(const BitPeek& br, uint64_t& p, uint32_t& b, uint32_t& z, uint32_t& nblocks,
                                            const uint64_t end, const uint64_t limit, const HuffDecTab* __restrict__ tabs,
                                            const uint16_t (*s_fast)[1 << kLutBits], int16_t* __restrict__ out, uint64_t blk,
                                            const uint64_t nblk, int* corrupt)
{
    while (p < end && p < limit) {
        const int cls = b >= 4;
        const int ti = (z == 0 ? 0 : 2) + cls;
        const uint32_t w = br.peek32(p);
        const uint32_t e = huff_lookup(tabs + ti, s_fast[ti], w);
        uint32_t len = e >> 8, sym = e & 255u;
        if (len == 0) {            // no such code: only legal while speculating
            if (kWrite && corrupt) *corrupt = 1;
            len = 1, sym = 0xffu;  // skip a bit, keep the state
            p += 1;
            continue;
        }
        const uint32_t s = (z == 0) ? (sym & 15u) : (sym & 15u);
        const uint32_t run = (z == 0) ? 0u : (sym >> 4);
        if (z != 0 && sym == 0) {  // EOB
            p += len;
            z = 64;
        } else {
            // value bits follow the code; len + s <= 32 holds for s <= 16
            uint32_t vbits = 0;
            if (s) {
                const uint32_t w2 = (len + s <= 32) ? (w << len) : br.peek32(p + len);
                vbits = w2 >> (32 - s);
            }
            p += len + s;
            z += run;
            if (z > 63) {          // run past the end of the block (src/decoder/jpezy_decoder.hpp:619)
                if (kWrite && corrupt) *corrupt = 1;
                z = 64;
            } else {
                if (kWrite && blk < nblk) {
                    int v = int(vbits);
                    if (s && !(vbits & (1u << (s - 1)))) v -= (1 << s) - 1;
                    if (s) out[blk * 64 + z] = int16_t(v);
                }
                z += 1;
            }
        }
        if (z >= 64) {
            z = 0;
            b = (b == 5) ? 0 : b + 1;
            ++nblocks;
            ++blk;
            if (kWrite && blk >= nblk) return;
        }
    }
}

__device__ __forceinline__ void load_dec_tabs(const HuffDecTab* __restrict__ tabs, uint16_t (*s_fast)[1 << kLutBits])
{
    for (int i = threadIdx.x; i < 4 * (1 << kLutBits); i += blockDim.x) s_fast[i >> kLutBits][i & ((1 << kLutBits) - 1)] = tabs[i >> kLutBits].fast[i & ((1 << kLutBits) - 1)];
}

// ---- D0: un-stuffing -------------------------------------------------------------------------------
// keep[i] = !(byte[i] == 0x00 && byte[i-1] == 0xFF)
__device__ __forceinline__ uint32_t keep_mask16(const uint8_t* __restrict__ src, uint64_t i0, uint64_t n, uint8_t* bytes)
{
    uint32_t m = 0;
    uint8_t prev = i0 ? src[i0 - 1] : 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint64_t idx = i0 + i;
        const uint8_t c = idx < n ? src[idx] : 0;
        bytes[i] = c;
        if (idx < n && !(c == 0 && prev == 0xff)) m |= 1u << i;
        prev = c;
    }
    return m;
}

__device__ __forceinline__ uint32_t cta_scan_excl(uint32_t v, uint32_t* s_warp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        const uint32_t s = s_warp[w];
        if (w < wid) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + x - v;
}

__global__ void __launch_bounds__(kDecThreads) k_unstuff_count(const DecParams p)
{
    __shared__ uint32_t s_warp[kDecThreads / 32];
    const size_t img = blockIdx.y;
    const uint64_t n = p.scan_bytes[img];
    const uint8_t* src = p.scan + img * p.slot;
    const uint32_t nch = uint32_t((n + 4095) / 4096);
    for (uint32_t ch = blockIdx.x; ch < nch; ch += gridDim.x) {
        uint8_t bytes[16];
        const uint64_t i0 = (uint64_t(ch) * kDecThreads + threadIdx.x) * 16;
        const uint32_t m = i0 < n ? keep_mask16(src, i0, n, bytes) : 0u;
        uint32_t total;
        cta_scan_excl(__popc(m), s_warp, &total);
        if (threadIdx.x == 0) p.chunk_cnt[img * p.nchunk + ch] = total;
    }
}

// generic per-image exclusive scan of uint32 counts with a 64-bit carry; one CTA (1024 threads) per image
__global__ void __launch_bounds__(1024) k_scan_chunks(const uint32_t* __restrict__ cnt, uint64_t* __restrict__ base, uint32_t stride,
                                                      const uint64_t* __restrict__ nbytes, uint32_t unit, uint64_t* __restrict__ total_out)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint64_t s_carry;
    const size_t img = blockIdx.x;
    const uint32_t n = uint32_t((nbytes[img] + unit - 1) / unit);
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n; c0 += 1024) {
        const uint32_t c = c0 + threadIdx.x;
        const uint32_t v = c < n ? cnt[img * stride + c] : 0u;
        uint32_t total;
        const uint32_t off = cta_scan_excl(v, s_warp, &total);
        const uint64_t carry = s_carry;
        if (c < n) base[img * stride + c] = carry + off;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) total_out[img] = s_carry;
}

__global__ void __launch_bounds__(kDecThreads) k_unstuff_write(const DecParams p)
{
    __shared__ uint32_t s_warp[kDecThreads / 32];
    const size_t img = blockIdx.y;
    const uint64_t n = p.scan_bytes[img];
    const uint8_t* src = p.scan + img * p.slot;
    uint8_t* dst = p.ustream + img * p.uslot;
    const uint32_t nch = uint32_t((n + 4095) / 4096);
    for (uint32_t ch = blockIdx.x; ch < nch; ch += gridDim.x) {
        uint8_t bytes[16];
        const uint64_t i0 = (uint64_t(ch) * kDecThreads + threadIdx.x) * 16;
        const uint32_t m = i0 < n ? keep_mask16(src, i0, n, bytes) : 0u;
        uint32_t total;
        const uint32_t off = cta_scan_excl(__popc(m), s_warp, &total);
        uint64_t o = p.chunk_base[img * p.nchunk + ch] + off;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (m & (1u << i)) dst[o++] = bytes[i];
    }
    // zero slack after the data so that peeks past the end read zeros
    if (blockIdx.x == 0 && threadIdx.x < 64) {
        const uint64_t u = p.ubytes[img];
        if (u + threadIdx.x < p.uslot) dst[u + threadIdx.x] = 0;
    }
}

// ---- D1a: speculative decode / synchronisation rounds -------------------------------------------------
// round 0: every subsequence starts at its own first bit in state (b=0, z=0).
// round k: subsequence i restarts from the end state of subsequence i-1 when that changed.
__global__ void __launch_bounds__(kDecThreads) k_sync_decode(const DecParams p, const int round)
{
    __shared__ uint16_t s_fast[4][1 << kLutBits];
    load_dec_tabs(p.tabs, s_fast);
    __syncthreads();
    const size_t img = blockIdx.y;
    const uint64_t total_bits = p.ubytes[img] * 8;
    const uint32_t i = blockIdx.x * kDecThreads + threadIdx.x;
    const uint64_t start = uint64_t(i) * p.sub_bits;
    if (start >= total_bits || i >= p.nsub) return;
    const uint32_t* sin = (round & 1) ? p.state_a : p.state_b;   // round 0 writes a, round 1 reads a writes b, ...
    uint32_t* sout = (round & 1) ? p.state_b : p.state_a;
    const uint8_t* din = (round & 1) ? p.dirty_a : p.dirty_b;
    uint8_t* dout = (round & 1) ? p.dirty_b : p.dirty_a;
    const size_t si = img * p.nsub + i;
    const uint64_t end = start + p.sub_bits;
    BitPeek br{p.ustream + img * p.uslot};
    uint64_t pos;
    uint32_t b, z, n = 0;
    if (round == 0) {
        pos = start, b = 0, z = 0;
    } else {
        if (i == 0 || !din[si - 1]) {   // predecessor unchanged: keep my end state
            sout[si] = sin[si];
            dout[si] = 0;
            return;
        }
        const uint32_t ps = sin[si - 1];
        pos = start + (ps & 63u), b = (ps >> 6) & 7u, z = (ps >> 9) & 63u;
    }
    decode_span<false>(br, pos, b, z, n, end, total_bits, p.tabs, s_fast, nullptr, 0, 0, nullptr);
    const uint64_t over = pos > end ? pos - end : 0;
    const uint32_t st = pack_state(uint32_t(over), b, z, n);
    sout[si] = st;
    if (round == 0) {
        dout[si] = 1;
    } else {
        const bool ch = ((st ^ sin[si]) & kStateSyncMask) != 0;
        dout[si] = ch ? 1 : 0;
        if (ch) atomicAdd(p.changed, 1ull);
    }
}

// block counts of the converged states -> sub_blk (exclusive), and per-image status
__global__ void __launch_bounds__(1024) k_scan_blocks(const DecParams p, const int final_round)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const size_t img = blockIdx.x;
    const uint32_t* st = (final_round & 1) ? p.state_b : p.state_a;   // buffer written by the last round
    const uint64_t total_bits = p.ubytes[img] * 8;
    const uint32_t n = uint32_t(min(uint64_t(p.nsub), (total_bits + p.sub_bits - 1) / p.sub_bits));
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n; c0 += 1024) {
        const uint32_t c = c0 + threadIdx.x;
        const uint32_t v = c < n ? (st[img * p.nsub + c] >> 15) & 4095u : 0u;
        uint32_t total;
        const uint32_t off = cta_scan_excl(v, s_warp, &total);
        const uint32_t carry = s_carry;
        if (c < n) p.sub_blk[img * p.nsub + c] = carry + off;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && p.status) p.status[img] = s_carry >= p.nblk ? 0 : JPEZYB200_ECORRUPT;
}

// ---- D1c: final pass, writes the non-zero coefficients (buffer pre-zeroed) ------------------------------
__global__ void __launch_bounds__(kDecThreads) k_write_coefs(const DecParams p, const int final_round)
{
    __shared__ uint16_t s_fast[4][1 << kLutBits];
    load_dec_tabs(p.tabs, s_fast);
    __syncthreads();
    const size_t img = blockIdx.y;
    const uint64_t total_bits = p.ubytes[img] * 8;
    const uint32_t i = blockIdx.x * kDecThreads + threadIdx.x;
    const uint64_t start = uint64_t(i) * p.sub_bits;
    if (start >= total_bits || i >= p.nsub) return;
    const uint32_t* st = (final_round & 1) ? p.state_b : p.state_a;
    const size_t si = img * p.nsub + i;
    uint64_t pos = start;
    uint32_t b = 0, z = 0, n = 0;
    if (i) {
        const uint32_t ps = st[si - 1];
        pos = start + (ps & 63u), b = (ps >> 6) & 7u, z = (ps >> 9) & 63u;
    }
    const uint64_t blk = p.sub_blk[si];
    if (blk >= p.nblk) return;
    BitPeek br{p.ustream + img * p.uslot};
    int corrupt = 0;
    decode_span<true>(br, pos, b, z, n, start + p.sub_bits, total_bits, p.tabs, s_fast, p.coefs + img * p.coef_stride, blk, p.nblk,
                      &corrupt);
    if (corrupt && p.status) p.status[img] = JPEZYB200_ECORRUPT;
}

// ---- D1d: DC differences -> absolute DC (pred_dct[sc] += diff, src/decoder/jpezy_decoder.hpp:596-597) -----
// tile = 256 MCUs; per tile the sums of the Y (4 blocks), Cb, Cr differences
__global__ void __launch_bounds__(256) k_dc_sum(const DecParams p)
{
    __shared__ int s_red[3][8];
    const size_t img = blockIdx.y;
    const uint32_t m = blockIdx.x * 256 + threadIdx.x;
    int y = 0, cb = 0, cr = 0;
    if (m < p.nmcu) {
        const int16_t* c = p.coefs + img * p.coef_stride + size_t(m) * 384;
        y = int(c[0]) + int(c[64]) + int(c[128]) + int(c[192]);
        cb = c[256], cr = c[320];
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        y += __shfl_xor_sync(0xffffffffu, y, d);
        cb += __shfl_xor_sync(0xffffffffu, cb, d);
        cr += __shfl_xor_sync(0xffffffffu, cr, d);
    }
    if ((threadIdx.x & 31) == 0) s_red[0][threadIdx.x >> 5] = y, s_red[1][threadIdx.x >> 5] = cb, s_red[2][threadIdx.x >> 5] = cr;
    __syncthreads();
    if (threadIdx.x < 3) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += s_red[threadIdx.x][w];
        p.dc_part[(img * p.ndc_tiles + blockIdx.x) * 3 + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(32) k_dc_scan_tiles(const DecParams p)
{
    // one warp per (image, component): sequential over tiles in chunks of 32 with a shuffle scan
    const size_t img = blockIdx.x;
    const int comp = blockIdx.y;
    int carry = 0;
    for (uint32_t t0 = 0; t0 < p.ndc_tiles; t0 += 32) {
        const uint32_t t = t0 + threadIdx.x;
        int v = t < p.ndc_tiles ? p.dc_part[(img * p.ndc_tiles + t) * 3 + comp] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if (int(threadIdx.x) >= d) x += y;
        }
        if (t < p.ndc_tiles) p.dc_part[(img * p.ndc_tiles + t) * 3 + comp] = carry + x - v;   // exclusive
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
}

__global__ void __launch_bounds__(256) k_dc_apply(const DecParams p)
{
    __shared__ int s_w[3][8];
    const size_t img = blockIdx.y;
    const uint32_t m = blockIdx.x * 256 + threadIdx.x;
    int16_t* c = p.coefs + img * p.coef_stride + size_t(min(m, p.nmcu - 1)) * 384;
    int d[6] = {0, 0, 0, 0, 0, 0};
    if (m < p.nmcu) {
#pragma unroll
        for (int k = 0; k < 6; ++k) d[k] = c[k * 64];
    }
    int v[3] = {d[0] + d[1] + d[2] + d[3], d[4], d[5]};
    int incl[3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int x = v[k];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, s);
            if (lane >= s) x += y;
        }
        incl[k] = x;
        if (lane == 31) s_w[k][wid] = x;
    }
    __syncthreads();
    if (m >= p.nmcu) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int base = p.dc_part[(img * p.ndc_tiles + blockIdx.x) * 3 + k];
        for (int w = 0; w < wid; ++w) base += s_w[k][w];
        incl[k] += base;   // inclusive prefix up to and including this MCU
    }
    int py = incl[0] - v[0];   // predictor before this MCU
    py += d[0]; c[0] = int16_t(py);
    py += d[1]; c[64] = int16_t(py);
    py += d[2]; c[128] = int16_t(py);
    py += d[3]; c[192] = int16_t(py);
    c[256] = int16_t(incl[1]);
    c[320] = int16_t(incl[2]);
}

}  // namespace jz
