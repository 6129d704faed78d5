// enc_transform2.cuh -- stages E1+E2, second generation: planar RGB -> quantised zig-zag int16 coefficients.
//
// Replaces make_YCC (src/encoder/jpezy_encoder.hpp:90-144), RGB::Y/Cb/Cr (:244-256), DCT (:146-166) and quantization
// (:168-172) of the reference, like enc_transform.cuh, whose kernels stay as the path for images whose rows are not
// 16-byte aligned and as the A/B variants.  What is different here (ncu of round 1: the kernel was bound by instruction
// issue, 66 thread-instruction slots per pixel, a third of them colour conversion):
//
//  * the pixel rows of a tile (T MCUs of one MCU row, 3 planes x 16 rows) are fetched by the TMA unit (cp.async.bulk, one
//    bulk copy per row and plane, completion on an mbarrier) and the tile's coefficients leave through one bulk store from
//    the staging buffer: no per-thread global loads, stores or address arithmetic on the hot path;
//  * everything floating point is packed f32x2 (FADD2 / FMUL2 / FFMA2): one instruction works on the same sample of two
//    blocks -- the upper and the lower luma block of an 8-pixel column strip (Y0|Y2, Y1|Y3), or Cb|Cr -- so both blocks of a
//    pair share quantisation constants and no repacking is ever needed;
//  * eight lanes per block pair: lane y converts the two pixel rows y and y+8 (8 pixels each) and transforms them
//    along x; the 8x8 transpose goes through a padded, conflict-free per-warp scratch; lane j then transforms column j
//    and quantises its 8+8 coefficients.  A warp owns its four block pairs from the pixels to the coefficients: the only
//    CTA-wide barriers are the one behind the (rare) fix-up queue and the one in front of the bulk store;
//  * colour conversion without conversions: the weighted sums 299r+587g+114b-128000 (and the two chroma sums, in 1e-4 units)
//    are IDP.2A chains on the packed bytes that start from the bit pattern of 1.5*2^23, so the accumulator *is* the float
//    12582912 + sum; truncation toward zero of sum/1000 is one FFMA2.RZ against a sign-matched 2^23 (the product
//    sum * fl(0.001) is evaluated exactly inside the FMA and fl(0.001) > 0.001 never carries a non-multiple across an
//    integer: tools/colour_trunc_check.py enumerates every sum).  Exact multiples -- where the reference's own FP64
//    rounding decides -- are detected from the residual sum - 1000*Y (one product per thread) and patched from the
//    64 KiB table of k_build_y_exact (luma) or by evaluating the reference's FP64 expression (chroma);
//  * the DC coefficient needs no special path: the sum of the 64 samples is exact in FP32 and the reference's
//    int(int(((S*c)*c)/4)/q) equals trunc(S * (1-2^-20)/(8q)) for every possible S (tools/colour_trunc_check.py);
//  * a block pair whose scaled coefficients are all below 1 - 2G (most pairs of photographic content) is finished after
//    one 3-input-max chain and a vote: the staging buffer is pre-zeroed, only the two DC values are stored.
//
// Numerics are those of enc_transform.cuh: same AAN flowgraph, same guard band G (tools/aan_error_bound.py; the samples
// are now level-shifted integers in [-128, 127] instead of [0, 255], the bound covers both), FP64 re-evaluation of the
// flagged coefficients from the staged pixels, the reference's exact operation order within 1e-9 of a quantiser multiple.
// Algorithmic HBM traffic: 3 B/px read + 3 B/px written = 6 B/px.
#pragma once
#include "enc_transform.cuh"

namespace jz {

// ---- async proxy: bulk copies and their barrier -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the barrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ f32x2 fma2_rz(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float(uint32_t(v)); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float(uint32_t(v >> 32)); }
__device__ __forceinline__ float fmax3_abs(float a, float b, float c)
{
    float r;
    asm("{\n\t.reg .f32 x, y;\n\tabs.f32 x, %1;\n\tabs.f32 y, %2;\n\tmax.f32 %0, x, y, %3;\n\t}" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---- constants of the second-generation kernel --------------------------------------------------------------------
// per (class, column j) the eight multipliers of coefficients (0..7, j) as (K, K) pairs, the guard bands, the zig-zag positions
struct Q2Tab {
    float2 K[2][8][8];     // [class][j][i]: (K, K) of coefficient (i, j); K of the DC coefficient carries the factor 1 - 2^-20
    float G[2][8][8];      // [class][j][i]: guard band in w units
    uint2 izz[8];          // [j]: zig-zag positions of coefficients (0..7, j), one byte each
    uint32_t thr[2];       // bit pattern of 1 - 2 Gmax: |w| below it quantises to 0 whatever the coefficient
    uint32_t pad[2];
};
static __device__ Q2Tab gQ2;

constexpr float kMagic15 = 12582912.0f;    // 1.5 * 2^23: float(kMagic15 + n) has the integer n in its low mantissa bits (|n| < 2^22)
constexpr uint32_t kMagic15Bits = 0x4B400000u;

// accumulator + c * byte E of w (dp2a: two signed 16-bit factors against two unsigned bytes; E picks the half and the slot)
template <int E>
__device__ __forceinline__ uint32_t mac_byte(int c, uint32_t w, uint32_t acc)
{
    const int cc = (E & 1) ? int(uint32_t(c) << 16) : int(uint32_t(c) & 0xffffu);
    return uint32_t((E & 2) ? dp2a_hi_su(cc, w, int(acc)) : dp2a_lo_su(cc, w, int(acc)));
}

// trunc(n / D) for the two integers held (as floats) in n2, D = 1000 or 10000, RCP = fl(1/D) rounded up; also the residual
// n - D * trunc(n / D), which is 0 exactly for the multiples of D
// (sgn = 0x80000000 and m23 = 0x4B000000 come in registers: with both as immediates the sign transfer is two LOP3, not one)
__device__ __forceinline__ f32x2 trunc_div2(f32x2 n2, float rcp, float negd, uint32_t sgn, uint32_t m23, f32x2& res)
{
    uint32_t mlo, mhi;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(mlo) : "r"(uint32_t(n2)), "r"(sgn), "r"(m23));            // (n & sgn) | m23
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(mhi) : "r"(uint32_t(n2 >> 32)), "r"(sgn), "r"(m23));
    const f32x2 m = (unsigned long long)mlo | ((unsigned long long)mhi << 32);      // +-2^23, the sign of n
    const f32x2 q = sub2(fma2_rz(n2, pk2(rcp, rcp), m), m);
    res = fma2(q, pk2(negd, negd), n2);
    return q;
}
constexpr float kRcp1000 = 0.001f;                 // 0x3a83126f > 1/1000
constexpr float kRcp10000 = 1.00000005e-4f;        // 0x38d1b718 = nextafter(fl(1e-4)) > 1/10000

// corrections of the exact luma cases of 8 pixels of one staged row: nibble x = ref_Y - exact (-1, 0 or 1)
__device__ __noinline__ uint32_t luma_fix_row(const uint8_t* pr, const uint8_t* pg, const uint8_t* pb, const int8_t* __restrict__ yx)
{
    uint32_t nib = 0;
#pragma unroll 1
    for (int x = 0; x < 8; ++x) {
        const int r = pr[x], g = pg[x], b = pb[x];
        if ((299 * r + 587 * g + 114 * b) % 1000 == 0) nib |= (uint32_t(int(yx[r | (g << 8)])) & 15u) << (4 * x);
    }
    return nib;
}
// the same for the 8 chroma samples (pixels 0, 2, .., 14) of one staged row: low word Cb, high word Cr
__device__ __noinline__ unsigned long long chroma_fix_row(const uint8_t* pr, const uint8_t* pg, const uint8_t* pb)
{
    uint32_t nb = 0, nr = 0;
#pragma unroll 1
    for (int x = 0; x < 8; ++x) {
        const int r = pr[2 * x], g = pg[2 * x], b = pb[2 * x];
        const int ncb = -1687 * r - 3313 * g + 5000 * b, ncr = 5000 * r - 4187 * g - 813 * b;
        if (ncb % 10000 == 0) nb |= (uint32_t(ref_Cb(r, g, b) - ncb / 10000) & 15u) << (4 * x);
        if (ncr % 10000 == 0) nr |= (uint32_t(ref_Cr(r, g, b) - ncr / 10000) & 15u) << (4 * x);
    }
    return (unsigned long long)nb | ((unsigned long long)nr << 32);
}
__device__ __forceinline__ float nibble_f(uint32_t nib, int x) { return float(int(nib << (28 - 4 * x)) >> 28); }

// exact sample (level-shifted Y, Cb or Cr) at (y, x) of block `blk` (scan order within the tile) from the staged pixels
template <int ROW>
__device__ __forceinline__ int tile_sample(const uint8_t* s_in, uint32_t blk, int y, int x, int hlast, int gray)
{
    const uint32_t m = blk / 6u, k = blk - m * 6u;
    int py, px;
    if (k < 4u) py = int(k >> 1) * 8 + y, px = int(m) * 16 + int(k & 1u) * 8 + x;
    else py = 2 * y, px = int(m) * 16 + 2 * x;
    py = min(py, hlast);
    const int r = s_in[py * ROW + px], g = s_in[(16 + py) * ROW + px], b = s_in[(32 + py) * ROW + px];
    if (k < 4u) return fast_Y(r, g, b);
    if (gray) return 0;
    return k == 4u ? fast_Cb(r, g, b) : fast_Cr(r, g, b);
}

// tier 2 / tier 3 decision on the FP64 separable sum `acc` of coefficient (i, j) of block blk (requant_finish of the
// first-generation kernel, samples re-derived from the pixels)
template <int ROW>
__device__ __noinline__ int requant_finish2(double acc, const uint8_t* s_in, uint32_t blk, int i, int j, int q, int hlast, int gray, unsigned long long* counter)
{
    double v = acc * 0.25 * (i ? 1.0 : 0.70710678118654752440) * (j ? 1.0 : 0.70710678118654752440);
    const double k = rint(v / double(q));
    if (k != 0.0 && fabs(v - k * double(q)) < 1e-9) {
        double sum = 0.0;
        for (int y = 0; y < 8; ++y) {
            const double ci = cC.cos_ref[i * 8 + y];
            for (int x = 0; x < 8; ++x)
                sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(double(tile_sample<ROW>(s_in, blk, y, x, hlast, gray)), cC.cos_ref[j * 8 + x]), ci));
        }
        const double cu = j ? 1.0 : cC.inv_sqrt2_ref, cv = i ? 1.0 : cC.inv_sqrt2_ref;
        v = __dmul_rn(__dmul_rn(__dmul_rn(sum, cu), cv), 0.25);
        atomicAdd(counter, 1ull);
    }
    return __double2int_rz(v) / q;
}

template <int T, int NST>
struct Fwd2 {
    static constexpr int kThreads = T * 24;           // one lane per (block pair, row / column): 3 T pairs x 8
    static constexpr int kRow = T * 16 + 16;          // bytes per staged pixel row; == 16 (mod 128): 8-byte reads down a column hit distinct banks
    static constexpr int kIn = 48 * kRow;             // one input stage: 3 planes x 16 rows
    static constexpr int kPairRow = 80;               // 8 x f32x2 + 16: the lanes' 16-byte row stores land on distinct banks
    static constexpr int kPair = 8 * kPairRow + 64;   // == 64 (mod 128): the two pairs of a half warp read disjoint banks
    static constexpr int kMid = T * 3 * kPair;
    static constexpr int kOut = T * 768;              // one staging buffer (two of them: a bulk store may still be reading the other)
    static constexpr int kMeta = T * 6 * 4;
    static constexpr int kSmem = NST * kIn + kMid + 2 * kOut + 2 * kMeta + kFixCap * 2 + 16 + NST * 8 + int(sizeof(Q2Tab));
    static_assert(kRow % 128 == 16 && kIn % 16 == 0 && kMid % 16 == 0, "layout");
};

// Persistent CTAs: CTA b transforms tiles b, b + gridDim.x, ... (a tile = T MCUs of one MCU row of one image).  The pixel rows
// of tile i + NST - 1 are already in flight (TMA bulk copies into a ring of NST stages, one mbarrier per stage) while tile i is
// transformed; the coefficients of tile i - 1 leave through a bulk store while tile i fills the other staging buffer.  One
// CTA-wide barrier per tile (two when the fix-up queue is not empty).
template <int T, int NST>
__global__ void __launch_bounds__(T * 24, T == 8 ? 4 : 2) k_fwd_transform2(const FwdParams p, const uint32_t ntiles, const uint32_t tiles_per_row)
{
    using C = Fwd2<T, NST>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_in0 = smem;                                           // [NST][3 planes][16 rows][kRow]
    uint8_t* s_mid = s_in0 + NST * C::kIn;                           // [3T pairs][8][kPairRow]: row-transformed samples, (A, B) packed
    uint8_t* s_out0 = s_mid + C::kMid;                               // [2][6T blocks][64] int16 zig-zag coefficients, scan order
    Q2Tab* s_tab = reinterpret_cast<Q2Tab*>(s_out0 + 2 * C::kOut);   // quantisation constants (copied once per CTA), 16-byte aligned
    uint32_t* s_meta0 = reinterpret_cast<uint32_t*>(s_tab + 1);
    uint16_t* s_fix = reinterpret_cast<uint16_t*>(s_meta0 + 2 * T * 6);
    uint32_t* s_nfix = reinterpret_cast<uint32_t*>(s_fix + kFixCap);
    const uint32_t bar0 = smem_u32(s_nfix + 4);                      // NST barriers of 8 bytes; s_nfix[0..2]: queue lengths of tiles it % 3
    static_assert(sizeof(Q2Tab) % 16 == 0 && (2 * T * 6 * 4) % 8 == 0, "alignment of the barriers");

    const int t = threadIdx.x, lane = t & 31;
    const uint32_t tiles_per_img = tiles_per_row * p.VU;
    // tile id -> image, MCU row, first MCU
    auto geom = [&](uint32_t id, uint32_t& img, uint32_t& my, uint32_t& mx0) {
        img = id / tiles_per_img;
        const uint32_t rem = id - img * tiles_per_img;
        my = rem / tiles_per_row;
        mx0 = (rem - my * tiles_per_row) * T;
    };
    // warp 0: the bulk copies of tile `id` into stage `st` (one per pixel row and plane)
    auto issue = [&](uint32_t id, int st) {
        uint32_t img, my, mx0;
        geom(id, img, my, mx0);
        const uint32_t nvalid = min(uint32_t(T), p.HU - mx0), gy0 = (p.row0 + my) * 16u;
        const int hlast = int(min(15u, p.H - 1u - gy0));
        const uint32_t rowbytes = nvalid * 16u, bar = bar0 + 8u * st;
        if (lane == 0) mbar_expect_tx(bar, 3u * uint32_t(hlast + 1) * rowbytes);
        __syncwarp();
        for (int c = lane; c < 48; c += 32) {
            const int plane = c >> 4, row = c & 15;
            if (row <= hlast) {
                const uint8_t* src = (plane == 0 ? p.r : (plane == 1 ? p.g : p.b)) + size_t(img) * p.plane_stride + size_t(gy0 - p.y_origin + row) * p.W + size_t(mx0) * 16u;
                bulk_g2s(smem_u32(s_in0 + st * C::kIn + c * C::kRow), src, rowbytes, bar);
            }
        }
    };

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(bar0 + 8u * s, 1);
        s_nfix[0] = s_nfix[1] = s_nfix[2] = 0;
        mbar_fence_init();
    }
    for (uint32_t i = t; i < sizeof(Q2Tab) / 4; i += C::kThreads) reinterpret_cast<uint32_t*>(s_tab)[i] = reinterpret_cast<const uint32_t*>(&gQ2)[i];
    __syncthreads();
    pdl_wait();
    if (t < 32) {
#pragma unroll
        for (int s = 0; s < NST - 1; ++s) {
            const uint32_t id = blockIdx.x + s * gridDim.x;
            if (id < ntiles) issue(id, s);
        }
    }

    // ---- the items of this thread: block pair pr, row (phase 1) / column (phase 2) sub ----
    const uint32_t pr = uint32_t(t) >> 3, sub = uint32_t(t) & 7u;
    const bool luma = pr < 2u * T;                                   // warp uniform (4 pairs per warp)
    const uint32_t mcu = luma ? (pr >> 1) : pr - 2u * T;
    const uint32_t blkA = luma ? mcu * 6u + (pr & 1u) : mcu * 6u + 4u, blkB = luma ? blkA + 2u : blkA + 1u;
    uint8_t* mid = s_mid + pr * C::kPair;
    const bool work = luma || !p.gray;                               // --gray: Cb = Cr = 0 (:61-64), the chroma blocks are all zero

    uint32_t sgn_bit, m23_bits;          // kept out of the immediate fields (trunc_div2)
    asm volatile("mov.u32 %0, 0x80000000;" : "=r"(sgn_bit));
    asm volatile("mov.u32 %0, 0x4B000000;" : "=r"(m23_bits));

    uint32_t it = 0;
    for (uint32_t id = blockIdx.x; id < ntiles; id += gridDim.x, ++it) {
        const int st = int(it % NST), ob = int(it & 1u);
        if (t < 32) {
            const uint32_t nid = id + (NST - 1) * gridDim.x;        // its stage was last read by tile it - 1 (behind the barrier below)
            if (nid < ntiles) issue(nid, int((it + NST - 1) % NST));
        }
        uint32_t img, my, mx0;
        geom(id, img, my, mx0);
        const uint32_t nvalid = min(uint32_t(T), p.HU - mx0);
        const int hlast = int(min(15u, p.H - 1u - (p.row0 + my) * 16u));   // last staged row; rows below replicate it (:101)
        const bool valid = mcu < nvalid;
        const uint8_t* s_in = s_in0 + st * C::kIn;
        int16_t* s_out = reinterpret_cast<int16_t*>(s_out0 + ob * C::kOut);
        uint32_t* s_meta = s_meta0 + ob * T * 6;
        uint32_t* nfix_p = s_nfix + it % 3u;
        {   // this lane's share of the pair's staging area starts as zeros: only non-zero coefficients are stored (the bulk store
            // that last read this buffer, two tiles ago, was waited for in front of the previous tile's barrier)
            reinterpret_cast<uint4*>(s_out + blkA * 64u)[sub] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(s_out + blkB * 64u)[sub] = make_uint4(0, 0, 0, 0);
        }
        mbar_wait(bar0 + 8u * st, (it / NST) & 1u);

        // ---- phase 1: colour conversion of rows sub and sub + 8 (luma) / row 2 sub (chroma), transform along x ----
        if (work) {
            f32x2 y2[8];
            if (luma) {
                const int rowA = min(int(sub), hlast), rowB = min(int(sub) + 8, hlast);
                const uint32_t col = mcu * 16u + (pr & 1u) * 8u;
                const uint8_t *pa = s_in + rowA * C::kRow + col, *pb = s_in + rowB * C::kRow + col;
                const uint2 ra = *reinterpret_cast<const uint2*>(pa), ga = *reinterpret_cast<const uint2*>(pa + 16 * C::kRow), ba = *reinterpret_cast<const uint2*>(pa + 32 * C::kRow);
                const uint2 rb = *reinterpret_cast<const uint2*>(pb), gb = *reinterpret_cast<const uint2*>(pb + 16 * C::kRow), bb = *reinterpret_cast<const uint2*>(pb + 32 * C::kRow);
                constexpr uint32_t kInit = kMagic15Bits - 128000u;
                f32x2 res[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    uint32_t fa, fb;
#define JZ_Y_SUM(E, R, G, B) mac_byte<E>(299, R, mac_byte<E>(587, G, mac_byte<E>(114, B, kInit)))
                    switch (x & 3) {
                    case 0: fa = JZ_Y_SUM(0, x < 4 ? ra.x : ra.y, x < 4 ? ga.x : ga.y, x < 4 ? ba.x : ba.y), fb = JZ_Y_SUM(0, x < 4 ? rb.x : rb.y, x < 4 ? gb.x : gb.y, x < 4 ? bb.x : bb.y); break;
                    case 1: fa = JZ_Y_SUM(1, x < 4 ? ra.x : ra.y, x < 4 ? ga.x : ga.y, x < 4 ? ba.x : ba.y), fb = JZ_Y_SUM(1, x < 4 ? rb.x : rb.y, x < 4 ? gb.x : gb.y, x < 4 ? bb.x : bb.y); break;
                    case 2: fa = JZ_Y_SUM(2, x < 4 ? ra.x : ra.y, x < 4 ? ga.x : ga.y, x < 4 ? ba.x : ba.y), fb = JZ_Y_SUM(2, x < 4 ? rb.x : rb.y, x < 4 ? gb.x : gb.y, x < 4 ? bb.x : bb.y); break;
                    default: fa = JZ_Y_SUM(3, x < 4 ? ra.x : ra.y, x < 4 ? ga.x : ga.y, x < 4 ? ba.x : ba.y), fb = JZ_Y_SUM(3, x < 4 ? rb.x : rb.y, x < 4 ? gb.x : gb.y, x < 4 ? bb.x : bb.y); break;
                    }
#undef JZ_Y_SUM
                    const f32x2 n2 = sub2((unsigned long long)fa | ((unsigned long long)fb << 32), pk2(kMagic15, kMagic15));
                    y2[x] = trunc_div2(n2, kRcp1000, -1000.0f, sgn_bit, m23_bits, res[x]);
                }
                // a weighted sum that is an exact multiple of 1000 (one pixel in a thousand): the reference's own FP64 rounding
                // decides, the 64 KiB table of k_build_y_exact holds its verdict per (r, g).  One product per lane finds out
                // whether the warp has such a pixel at all (4 warps in 10 do), one vote per column where.
                f32x2 prod = mul2(mul2(mul2(res[0], res[1]), mul2(res[2], res[3])), mul2(mul2(res[4], res[5]), mul2(res[6], res[7])));
                if (__any_sync(0xffffffffu, lo2(prod) * hi2(prod) == 0.0f)) {
#pragma unroll
                    for (int x = 0; x < 8; ++x) {
                        const bool za = lo2(res[x]) == 0.0f, zb = hi2(res[x]) == 0.0f;
                        if (__any_sync(0xffffffffu, za || zb)) {
                            float ca = 0.0f, cb = 0.0f;       // (the pixel's r and g come from the staged rows again: the words are dead by now)
                            if (za) ca = float(int(__ldg(p.y_exact + (uint32_t(pa[x]) | (uint32_t(pa[16 * C::kRow + x]) << 8)))));
                            if (zb) cb = float(int(__ldg(p.y_exact + (uint32_t(pb[x]) | (uint32_t(pb[16 * C::kRow + x]) << 8)))));
                            y2[x] = add2(y2[x], pk2(ca, cb));
                        }
                    }
                }
            } else {
                // chroma: the exact cases (r == g and b - r even, ...) are a few per cent of natural samples, so the formulas are
                // evaluated the way the reference does, in FP64 (half rate on this part) -- 8 samples per lane, a quarter of the pixels
                const int row = min(2 * int(sub), hlast);
                const uint8_t* pa = s_in + row * C::kRow + mcu * 16u;
                const uint4 rv = *reinterpret_cast<const uint4*>(pa), gv = *reinterpret_cast<const uint4*>(pa + 16 * C::kRow), bv = *reinterpret_cast<const uint4*>(pa + 32 * C::kRow);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int x = 0; x < 8; ++x) {          // decimation, not averaging (:116-143): the even pixels of the even rows
                    const int sh = (x & 1) * 16;
                    const double r = double((rw[x >> 1] >> sh) & 255u), g = double((gw[x >> 1] >> sh) & 255u), b = double((bw[x >> 1] >> sh) & 255u);
                    const double cb = __fma_rn(0.5000, b, __dsub_rn(-__dmul_rn(0.1687, r), __dmul_rn(0.3313, g)));            // 0.5 * b is exact
                    const double cr = __dsub_rn(__fma_rn(0.5000, r, -__dmul_rn(0.4187, g)), __dmul_rn(0.0813, b));            // 0.5 * r is exact
                    y2[x] = pk2(float(__double2int_rz(cb)), float(__double2int_rz(cr)));
                }
            }
            aan_fdct8_x2(y2[0], y2[1], y2[2], y2[3], y2[4], y2[5], y2[6], y2[7]);
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(mid + sub * C::kPairRow);
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[k] = make_ulonglong2(y2[2 * k], y2[2 * k + 1]);
        }
        __syncwarp();

        // ---- phase 2: transform along y, quantisation; lane sub holds column sub of both blocks ----
        if (work) {
            const uint32_t j = sub;
            const int cls = luma ? 0 : 1;
            f32x2 d[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = *reinterpret_cast<const f32x2*>(mid + i * C::kPairRow + j * 8);
            f32x2 w[8];
            {
                const ulonglong2* kq = reinterpret_cast<const ulonglong2*>(&s_tab->K[cls][j]);
                const ulonglong2 k0 = kq[0], k1 = kq[1], k2 = kq[2], k3 = kq[3];
                aan_fdct8_x2(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
                w[0] = mul2(d[0], k0.x), w[1] = mul2(d[1], k0.y), w[2] = mul2(d[2], k1.x), w[3] = mul2(d[3], k1.y);
                w[4] = mul2(d[4], k2.x), w[5] = mul2(d[5], k2.y), w[6] = mul2(d[6], k3.x), w[7] = mul2(d[7], k3.y);
            }
            int dcA = 0, dcB = 0;
            if (j == 0) {
                dcA = __float2int_rz(lo2(w[0])), dcB = __float2int_rz(hi2(w[0]));
                s_out[blkA * 64u] = int16_t(dcA), s_out[blkB * 64u] = int16_t(dcB);
            }
            // rows of coefficients in which some lane of the warp may quantise to a non-zero value: |w| >= 1 - 2 Gmax.  The
            // larger magnitude of the pair per row, the warp's maximum of its bit pattern (REDUX), one uniform compare.
            const uint32_t thr = s_tab->thr[cls];
            uint32_t gm = 0;        // non-zero zig-zag groups: bits 0..7 block A, 8..15 block B
            const uint2 izz = s_tab->izz[j];
            const float* G = s_tab->G[cls][j];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a = fmaxf(fabsf(lo2(w[i])), fabsf(hi2(w[i])));
                if (i == 0) a = j ? a : 0.0f;                                   // the DC coefficient went its own way
                if (__reduce_max_sync(0xffffffffu, valid ? __float_as_uint(a) : 0u) < thr) continue;
                const uint32_t zz = __byte_perm(i < 4 ? izz.x : izz.y, 0, 0x4440 + (i & 3));
                const float g = G[i];
                const f32x2 mg = pk2(kMagic15, kMagic15);
                const f32x2 dl = sub2(w[i], sub2(add2(w[i], mg), mg));          // w - rint(w)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float wv = h ? hi2(w[i]) : lo2(w[i]);
                    const float dv = h ? hi2(dl) : lo2(dl);
                    const uint32_t blk = h ? blkB : blkA;
                    const int q = __float2int_rz(wv);
                    if ((i | j) != 0 && valid) {
                        if (fabsf(dv) < g && fabsf(wv) > 0.5f) push_fix(nfix_p, s_fix, (blk << 6) | uint32_t(i * 8) | j);
                        if (q != 0) {
                            s_out[blk * 64u + zz] = int16_t(q);
                            gm |= 1u << ((zz >> 3) + 8u * h);
                        }
                    }
                }
            }
            if (__any_sync(0xffffffffu, gm != 0u)) {
                gm |= __shfl_xor_sync(0xffffffffu, gm, 1);
                gm |= __shfl_xor_sync(0xffffffffu, gm, 2);
                gm |= __shfl_xor_sync(0xffffffffu, gm, 4);
            }
            if (j == 0) {
                s_meta[blkA] = (uint32_t(dcA) & 0xffffu) | ((gm & 0xffu) << 16);
                s_meta[blkB] = (uint32_t(dcB) & 0xffffu) | ((gm & 0xff00u) << 8);
            }
        } else if (sub == 0) {
            s_meta[blkA] = 0, s_meta[blkB] = 0;
        }
        fence_async_smem();      // the staging buffer is read by the async proxy (bulk store)
        if (t == 0) bulk_wait_read0();      // the previous tile's store has read the other staging buffer: the next tile may fill it
        __syncthreads();
        if (t == 0) s_nfix[(it + 2u) % 3u] = 0;      // the previous tile's queue length: read by everyone before this barrier, used again by tile it + 2

        // ---- phase 2b: dense FP64 re-evaluation of the guard-band queue ----
        {
            const uint32_t nfix = *nfix_p;
            if (nfix) {
                if (nfix > kFixCap) {
                    // queue overflow (adversarial content): every AC coefficient of the tile is re-evaluated
                    for (uint32_t e = t; e < nvalid * 6u * 64u; e += C::kThreads) {
                        const uint32_t blk = e >> 6, ij = e & 63u, i = ij >> 3, j = ij & 7u;
                        if (ij == 0 || (p.gray && blk % 6u >= 4u)) continue;
                        double acc = 0.0;
                        for (int y = 0; y < 8; ++y) {
                            double row = 0.0;
                            for (int x = 0; x < 8; ++x) row = fma(double(tile_sample<C::kRow>(s_in, blk, y, x, hlast, p.gray)), cC.cos_ref[j * 8 + x], row);
                            acc = fma(row, cC.cos_ref[i * 8 + y], acc);
                        }
                        const int q = cC.quant[(blk % 6u) >= 4u][ij];
                        const int v = requant_finish2<C::kRow>(acc, s_in, blk, int(i), int(j), q, hlast, p.gray, p.guard_counter);
                        s_out[blk * 64u + cC.izz[ij]] = int16_t(v);
                        if (v) atomicOr(&s_meta[blk], 1u << (16 + (cC.izz[ij] >> 3)));
                    }
                } else {
                    // eight lanes per entry: lane s evaluates row s of the separable sum, a 3-step butterfly adds the rows
                    const uint32_t s = uint32_t(t) & 7u;
                    for (uint32_t e0 = 0; e0 < nfix; e0 += C::kThreads / 8) {
                        const uint32_t e = e0 + (uint32_t(t) >> 3);
                        const bool actv = e < nfix;
                        const uint32_t ent = actv ? uint32_t(s_fix[e]) : 0u;
                        const uint32_t blk = ent >> 6, ij = ent & 63u, i = ij >> 3, j = ij & 7u;
                        const double* cj = &gCosRef[j * 8];
                        double row = 0.0;
#pragma unroll 1
                        for (int x = 0; x < 8; ++x) row = fma(double(tile_sample<C::kRow>(s_in, blk, int(s), x, hlast, p.gray)), cj[x], row);
                        double part = row * gCosRef[i * 8 + s];
                        part += __shfl_xor_sync(0xffffffffu, part, 1);
                        part += __shfl_xor_sync(0xffffffffu, part, 2);
                        part += __shfl_xor_sync(0xffffffffu, part, 4);
                        if (actv && s == 0) {
                            const int q = cC.quant[(blk % 6u) >= 4u][ij];
                            const int v = requant_finish2<C::kRow>(part, s_in, blk, int(i), int(j), q, hlast, p.gray, p.guard_counter);
                            s_out[blk * 64u + cC.izz[ij]] = int16_t(v);
                            if (v) atomicOr(&s_meta[blk], 1u << (16 + (cC.izz[ij] >> 3)));
                        }
                    }
                }
                fence_async_smem();
                __syncthreads();
            }
        }

        // ---- phase 3: one bulk store of the tile's coefficients (scan order), side information by plain stores ----
        const size_t mcu0 = size_t(my) * p.HU + mx0;
        if (t == 0) {
            bulk_s2g(p.coefs + size_t(img) * p.coef_stride + mcu0 * 384, smem_u32(s_out), nvalid * 768u);
            bulk_commit();
        }
        if (p.bmeta && uint32_t(t) < nvalid * 6u) p.bmeta[size_t(img) * (p.coef_stride >> 6) + mcu0 * 6 + t] = s_meta[t];
    }
    if (t == 0) bulk_wait_read0();
}

}  // namespace jz
