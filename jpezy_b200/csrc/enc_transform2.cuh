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
    float2 K[2][4][8][2];  // [class][i / 2][j][i % 2]: (K, K) of coefficient (i, j); K of the DC coefficient carries the factor 1 - 2^-20.
                           // The eight column lanes j of a pair read 16 bytes each, 128 contiguous bytes: no bank conflicts
    float G[2][8][8];      // [class][i][j]: guard band in w units (the lanes j read eight neighbouring words)
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

// ---- shared-memory accesses through 32-bit shared-space addresses (the generic-pointer forms cost 64-bit address arithmetic) ----
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ f32x2 lds64x(uint32_t a)
{
    f32x2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64x(uint32_t a, f32x2 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts128x(uint32_t a, f32x2 lo, f32x2 hi) { asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(a), "l"(lo), "l"(hi) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(short(v)) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// wait with a watchdog: a protocol error must trap, not hang the device
// (try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes instead of polling -- the polling
// loop of the first version was a third of all executed instructions)
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity)
{
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (!ok && ++spins > (1u << 16)) __trap();
    } while (!ok);
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// exact sample (level-shifted Y, Cb or Cr) at (y, x) of block `blk` (scan order within the tile) from the staged pixels
template <int ROW>
__device__ __forceinline__ int tile_sample_s(uint32_t s_in, uint32_t blk, int y, int x, int hlast, int gray)
{
    const uint32_t m = blk / 6u, k = blk - m * 6u;
    int py, px;
    if (k < 4u) py = int(k >> 1) * 8 + y, px = int(m) * 16 + int(k & 1u) * 8 + x;
    else py = 2 * y, px = int(m) * 16 + 2 * x;
    py = min(py, hlast);
    const uint32_t a = s_in + py * ROW + px;
    const int r = int(lds8(a)), g = int(lds8(a + 16 * ROW)), b = int(lds8(a + 32 * ROW));
    if (k < 4u) return fast_Y(r, g, b);
    if (gray) return 0;
    return k == 4u ? fast_Cb(r, g, b) : fast_Cr(r, g, b);
}
// tier 2 / tier 3 decision on the FP64 separable sum `acc` of coefficient (i, j) of block blk (requant_finish of the
// first-generation kernel, samples re-derived from the staged pixels)
template <int ROW>
__device__ __noinline__ int requant_finish_s(double acc, uint32_t s_in, uint32_t blk, int i, int j, int q, int hlast, int gray, unsigned long long* counter)
{
    double v = acc * 0.25 * (i ? 1.0 : 0.70710678118654752440) * (j ? 1.0 : 0.70710678118654752440);
    const double k = rint(v / double(q));
    if (k != 0.0 && fabs(v - k * double(q)) < 1e-9) {
        double sum = 0.0;
        for (int y = 0; y < 8; ++y) {
            const double ci = cC.cos_ref[i * 8 + y];
            for (int x = 0; x < 8; ++x)
                sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(double(tile_sample_s<ROW>(s_in, blk, y, x, hlast, gray)), cC.cos_ref[j * 8 + x]), ci));
        }
        const double cu = j ? 1.0 : cC.inv_sqrt2_ref, cv = i ? 1.0 : cC.inv_sqrt2_ref;
        v = __dmul_rn(__dmul_rn(__dmul_rn(sum, cu), cv), 0.25);
        atomicAdd(counter, 1ull);
    }
    return __double2int_rz(v) / q;
}

// per-warp fix-up queue: (block, natural index) entries of 16 bits, length in its own word
__device__ __noinline__ void push_fix_w(uint32_t a_cnt, uint32_t a_fix, uint32_t blk_ij)
{
    uint32_t idx;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(idx) : "r"(a_cnt) : "memory");
    if (idx < 96u) asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_fix + 2u * idx), "h"(short(blk_ij)) : "memory");
}

constexpr int kWarpFix = 96;       // fix-up queue entries per compute warp and tile (more: the warp re-evaluates all of its blocks)

template <int T, int NST>
struct Fwd2 {
    static constexpr int kWarps = T * 24 / 32;        // compute warps: one lane per (block pair, row / column), 4 pairs per warp
    static constexpr int kThreads = T * 24 + 32;      // + the DMA warp
    static constexpr int kRow = T * 16 + 16;          // bytes per staged pixel row; == 16 (mod 128): 8-byte reads down a column hit distinct banks
    static constexpr int kIn = 48 * kRow;             // one input stage: 3 planes x 16 rows
    static constexpr int kPairRow = 80;               // 8 x f32x2 + 16: the lanes' 16-byte row stores land on distinct banks
    static constexpr int kPair = 8 * kPairRow + 64;   // == 64 (mod 128): the two pairs of a half warp read disjoint banks
    static constexpr int kOut = T * 768;              // one staging buffer (two of them: a bulk store may still be reading the other)
    static constexpr uint32_t oIn = 0, oMid = oIn + NST * kIn, oOut = oMid + T * 3 * kPair, oTab = oOut + 2 * kOut;
    static constexpr uint32_t oFix = oTab + uint32_t(sizeof(Q2Tab)), oCnt = oFix + kWarps * kWarpFix * 2, oBar = oCnt + 64;
    static constexpr int kSmem = int(oBar) + (2 * NST + 4) * 8;
    static_assert(kRow % 128 == 16 && kIn % 16 == 0 && oOut % 16 == 0 && oTab % 16 == 0 && sizeof(Q2Tab) % 16 == 0 && oBar % 8 == 0 && kWarps <= 16, "layout");
};

// How the tile ids advance from one tile of a CTA to its next (id += gridDim.x), precomputed on the host: no divisions in the loop
struct TileStep {
    uint32_t tiles_per_row, dimg, dmy, dbx;
    uint32_t flags;      // experiments: bit 0 = the DMA warp issues a tile's bulk copies from one lane instead of 32; bit 1 = fixed warp roles
};

// Persistent, warp-specialised CTAs.  CTA b transforms tiles b, b + gridDim.x, ... (a tile = T MCUs of one MCU row of one image).
//  * the DMA warp (one lane) keeps NST - 1 tiles of pixel rows in flight (TMA bulk copies into a ring of NST stages, `in_full`
//    barriers) and sends every finished tile's coefficients off with one bulk store (`out_full` / `out_empty`);
//  * the compute warps never meet at a CTA barrier: a warp owns its four block pairs from the pixels to the coefficients and to its
//    own fix-up queue; it tells the DMA warp through `in_empty` / `out_full` when it is done with a stage / a staging buffer.
// WST (warp stores): every compute warp sends its own eight blocks off (its own two staging areas, its own bulk-store groups)
// instead of handing the tile to the DMA warp: no warp ever waits for the slowest warp of the tile on the output side, only for
// its own store of two tiles ago.
template <int T, int NST, bool WST>
__global__ void __launch_bounds__(T * 24 + 32, T == 8 ? 4 : 2) k_fwd_transform2(const FwdParams p, const uint32_t ntiles, const TileStep ts)
{
    using C = Fwd2<T, NST>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t sm = smem_u32(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(sm));       // (kept in a register: rematerialising the shared window base costs three uniform ops per use)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t bar_in_full = sm + C::oBar, bar_in_empty = bar_in_full + 8u * NST, bar_out_full = bar_in_empty + 8u * NST, bar_out_empty = bar_out_full + 16u;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(bar_in_full + 8u * s, 1), mbar_init(bar_in_empty + 8u * s, C::kWarps);
        mbar_init(bar_out_full, C::kWarps), mbar_init(bar_out_full + 8u, C::kWarps);
        mbar_init(bar_out_empty, 1), mbar_init(bar_out_empty + 8u, 1);
        mbar_fence_init();
    }
    if (t < 16) sts32(sm + C::oCnt + 4u * t, 0u);
    __syncthreads();
    pdl_wait();

    // tile id -> image, MCU row, tile of the row (divisions once per thread; the loops advance by ts)
    uint32_t img, my, bx;
    {
        const uint32_t per_img = ts.tiles_per_row * p.VU;
        img = blockIdx.x / per_img;
        const uint32_t rem = blockIdx.x - img * per_img;
        my = rem / ts.tiles_per_row, bx = rem - my * ts.tiles_per_row;
    }
    auto advance = [&](uint32_t& im, uint32_t& y, uint32_t& x) {
        x += ts.dbx;
        if (x >= ts.tiles_per_row) x -= ts.tiles_per_row, ++y;
        y += ts.dmy;
        if (y >= p.VU) y -= p.VU, ++im;
        im += ts.dimg;
    };

    if (warp == C::kWarps) {
        // ================= DMA warp =================
        // all 32 lanes issue the bulk copies of a tile (one per pixel row and plane: 48 requests in two rounds); lane 0 alone
        // arms the barrier and sends the finished tiles off
        uint32_t limg = img, lmy = my, lbx = bx;       // next tile to load
        uint32_t lid = blockIdx.x;
        auto issue = [&](int st) {
            const uint32_t mx0 = lbx * T, nvalid = min(uint32_t(T), p.HU - mx0), gy0 = (p.row0 + lmy) * 16u;
            const uint32_t nrow = min(16u, p.H - gy0), rowbytes = nvalid * 16u, bar = bar_in_full + 8u * st;
            if (lane == 0) mbar_expect_tx(bar, 3u * nrow * rowbytes);
            __syncwarp();
            const size_t off = size_t(limg) * p.plane_stride + size_t(gy0 - p.y_origin) * p.W + size_t(mx0) * 16u;
            const uint32_t dst = sm + C::oIn + st * C::kIn;
            if (ts.flags & 1u) {
                if (lane == 0)
#pragma unroll 1
                    for (uint32_t row = 0; row < nrow; ++row) {
                        bulk_g2s(dst + row * C::kRow, p.r + off + size_t(row) * p.W, rowbytes, bar);
                        bulk_g2s(dst + (16 + row) * C::kRow, p.g + off + size_t(row) * p.W, rowbytes, bar);
                        bulk_g2s(dst + (32 + row) * C::kRow, p.b + off + size_t(row) * p.W, rowbytes, bar);
                    }
            } else {
#pragma unroll
                for (int c = lane; c < 48; c += 32) {
                    const uint32_t plane = uint32_t(c) >> 4, row = uint32_t(c) & 15u;
                    if (row < nrow) bulk_g2s(dst + uint32_t(c) * C::kRow, (plane == 0 ? p.r : (plane == 1 ? p.g : p.b)) + off + size_t(row) * p.W, rowbytes, bar);
                }
            }
            advance(limg, lmy, lbx);
            lid += gridDim.x;
        };
#pragma unroll 1
        for (int s = 0; s < NST - 1 && lid < ntiles; ++s) issue(s);
        uint32_t i = 0;
        int lst = NST - 1;          // stage of the next load
#pragma unroll 1
        for (uint32_t id = blockIdx.x; id < ntiles; id += gridDim.x, ++i) {
            if (lid < ntiles) {
                // tile i + NST - 1 goes where tile i - 1 was: wait until every compute warp is done with it
                if (i >= 1) mbar_wait_wd(bar_in_empty + 8u * lst, ((i - 1) / NST) & 1u);
                issue(lst);
                lst = lst + 1 == NST ? 0 : lst + 1;
            }
            if constexpr (WST) {
                if (lid >= ntiles) break;                        // nothing left to load: the compute warps store their tiles themselves
            } else {
                if (lane == 0) {
                    const uint32_t ob = i & 1u;
                    mbar_wait_wd(bar_out_full + 8u * ob, (i >> 1) & 1u);
                    const uint32_t mx0 = bx * T, nvalid = min(uint32_t(T), p.HU - mx0);
                    bulk_s2g(p.coefs + size_t(img) * p.coef_stride + (size_t(my) * p.HU + mx0) * 384, sm + C::oOut + ob * C::kOut, nvalid * 768u);
                    bulk_commit();
                    bulk_wait_read0();                           // the store has read the staging buffer (not: has reached memory)
                    mbar_arrive(bar_out_empty + 8u * ob);
                }
                __syncwarp();
                advance(img, my, bx);
            }
        }
        return;
    }

    // ================= compute warps =================
    // the quantisation constants come to shared memory while the first tile is in flight (a barrier of the compute warps only)
    for (uint32_t i = t; i < sizeof(Q2Tab) / 4; i += C::kWarps * 32) sts32(sm + C::oTab + 4u * i, reinterpret_cast<const uint32_t*>(&gQ2)[i]);
    asm volatile("bar.sync 1, %0;" ::"r"(C::kWarps * 32) : "memory");
    const uint32_t sub = uint32_t(t) & 7u;                           // row (phase 1) / column (phase 2) of the lane's block pair
    const uint32_t a_mid = sm + C::oMid + (uint32_t(t) >> 3) * C::kPair;
    const uint32_t a_fix = sm + C::oFix + warp * (kWarpFix * 2), a_cnt = sm + C::oCnt + 4u * warp;
    const uint32_t a_tab = sm + C::oTab;
    // Which four block pairs of the tile a warp takes changes from tile to tile: two thirds of the roles are luma pairs, one third
    // chroma pairs, and a chroma tile costs a warp more than a luma tile (FP64 colour conversion).  With fixed roles the luma warps
    // would wait for the chroma warps at every staging buffer (13 % of the warp samples of the first version); with six warps the
    // role advances by two per tile, so every warp sees luma, luma, chroma, ... and exactly two warps hold chroma roles at any tile.
    const uint32_t rot = (C::kWarps == 6 && !(ts.flags & 2u)) ? 2u : 0u;
    uint32_t role = uint32_t(warp);
    uint32_t sgn_bit, m23_bits;          // kept out of the immediate fields (trunc_div2)
    asm volatile("mov.u32 %0, 0x80000000;" : "=r"(sgn_bit));
    asm volatile("mov.u32 %0, 0x4B000000;" : "=r"(m23_bits));

    uint32_t st = 0, stpar = 0, i = 0;
#pragma unroll 1
    for (uint32_t id = blockIdx.x; id < ntiles; id += gridDim.x, ++i) {
        const uint32_t ob = i & 1u;
        const uint32_t pr = role * 4u + (uint32_t(lane) >> 3);            // block pair of this lane in this tile
        const bool luma = pr < 2u * T;                                   // warp uniform (4 pairs per warp)
        const uint32_t mcu = luma ? (pr >> 1) : pr - 2u * T;
        const uint32_t blkA = luma ? mcu * 6u + (pr & 1u) : mcu * 6u + 4u;
        const uint32_t dAB = luma ? 256u : 128u;                         // block B = Y2 / Y3 / Cr: bytes behind block A
        const bool work = luma || !p.gray;                               // --gray: Cb = Cr = 0 (:61-64), the chroma blocks are all zero
        const uint32_t cls = luma ? 0u : 1u;
        const uint32_t mx0 = bx * T, nvalid = min(uint32_t(T), p.HU - mx0);
        const int hlast = int(min(15u, p.H - 1u - (p.row0 + my) * 16u));   // last staged row; rows below replicate it (:101)
        const bool valid = mcu < nvalid;
        const uint32_t a_in = sm + C::oIn + st * C::kIn;
        // staging: WST -- the warp's own area, its blocks in the order they have in the coefficient array ([Y0 Y1 Y2 Y3] of its two
        // MCUs, [Cb Cr] of its four); otherwise the tile's area, shared by the warps
        const uint32_t a_wst = sm + C::oOut + uint32_t(warp) * 2048u + ob * 1024u;
        const uint32_t pl = uint32_t(lane) >> 3;
        const uint32_t a_outA = WST ? a_wst + (luma ? (pl >> 1) * 512u + (pl & 1u) * 128u : pl * 256u) : sm + C::oOut + ob * C::kOut + blkA * 128u;
        if constexpr (WST) {
            if (i >= 2 && lane == 0) bulk_wait_read1();      // this warp's stores of tile i - 2 have read the area
            __syncwarp();
        } else {
            if (i >= 2) mbar_wait_wd(bar_out_empty + 8u * ob, ((i >> 1) - 1u) & 1u);    // the store of tile i - 2 has read this buffer
        }
        // this lane's share of the pair's staging area starts as zeros: only non-zero coefficients are stored
        sts128(a_outA + sub * 16u, 0, 0, 0, 0);
        sts128(a_outA + dAB + sub * 16u, 0, 0, 0, 0);
        mbar_wait_wd(bar_in_full + 8u * st, stpar);

        // ---- phase 1: colour conversion of rows sub and sub + 8 (luma) / row 2 sub (chroma), transform along x ----
        if (work) {
            f32x2 y2[8];
            if (luma) {
                const uint32_t col = mcu * 16u + (pr & 1u) * 8u;
                const uint32_t pa = a_in + min(int(sub), hlast) * C::kRow + col, pb = a_in + min(int(sub) + 8, hlast) * C::kRow + col;
                const uint2 ra = lds64(pa), ga = lds64(pa + 16 * C::kRow), ba = lds64(pa + 32 * C::kRow);
                const uint2 rb = lds64(pb), gb = lds64(pb + 16 * C::kRow), bb = lds64(pb + 32 * C::kRow);
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
                const f32x2 prod = mul2(mul2(mul2(res[0], res[1]), mul2(res[2], res[3])), mul2(mul2(res[4], res[5]), mul2(res[6], res[7])));
                if (__any_sync(0xffffffffu, lo2(prod) * hi2(prod) == 0.0f)) {
#pragma unroll
                    for (int x = 0; x < 8; ++x) {
                        const bool za = lo2(res[x]) == 0.0f, zb = hi2(res[x]) == 0.0f;
                        if (__any_sync(0xffffffffu, za || zb)) {
                            float ca = 0.0f, cb = 0.0f;       // (the pixel's r and g come from the staged rows again: the words are dead by now)
                            if (za) ca = float(int(__ldg(p.y_exact + (lds8(pa + x) | (lds8(pa + 16 * C::kRow + x) << 8)))));
                            if (zb) cb = float(int(__ldg(p.y_exact + (lds8(pb + x) | (lds8(pb + 16 * C::kRow + x) << 8)))));
                            y2[x] = add2(y2[x], pk2(ca, cb));
                        }
                    }
                }
            } else {
                // chroma: the exact cases (r == g and b - r even, ...) are a few per cent of natural samples, so the formulas are
                // evaluated the way the reference does, in FP64 (half rate on this part) -- 8 samples per lane, a quarter of the pixels
                const uint32_t pa = a_in + min(2 * int(sub), hlast) * C::kRow + mcu * 16u;
                const uint4 rv = lds128(pa), gv = lds128(pa + 16 * C::kRow), bv = lds128(pa + 32 * C::kRow);
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
            const uint32_t dst = a_mid + sub * C::kPairRow;
#pragma unroll
            for (int k = 0; k < 8; ++k) sts64x(dst + 8 * k, y2[k]);       // (16-byte stores would need the pairs in adjacent registers)
        }
        __syncwarp();

        // ---- phase 2: transform along y, quantisation; lane sub holds column sub of both blocks ----
        int dcA = 0, dcB = 0;
        uint32_t gm = 0;        // non-zero zig-zag groups: bits 0..7 block A, 8..15 block B
        if (work) {
            const uint32_t j = sub;
            f32x2 w[8];
            {
                f32x2 d[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) d[k] = lds64x(a_mid + k * C::kPairRow + j * 8u);
                aan_fdct8_x2(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
                const uint32_t kq = a_tab + uint32_t(offsetof(Q2Tab, K)) + cls * 512u + j * 16u;
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    const uint4 kk = lds128(kq + 64 * k);
                    w[k] = mul2(d[k], (unsigned long long)kk.x | ((unsigned long long)kk.y << 32));
                    w[k + 1] = mul2(d[k + 1], (unsigned long long)kk.z | ((unsigned long long)kk.w << 32));
                }
            }
            if (j == 0) {
                dcA = __float2int_rz(lo2(w[0])), dcB = __float2int_rz(hi2(w[0]));
                sts16(a_outA, dcA), sts16(a_outA + dAB, dcB);
            }
            // rows of coefficients in which some lane of the warp may quantise to a non-zero value: |w| >= 1 - 2 Gmax.  The
            // larger magnitude of the pair per row, the warp's maximum of its bit pattern (REDUX), one uniform compare; the
            // five high rows (rarely alive in photographic content) share a first test.
            const uint32_t thr = lds32(a_tab + uint32_t(offsetof(Q2Tab, thr)) + 4u * cls);
            const uint2 izz = lds64(a_tab + uint32_t(offsetof(Q2Tab, izz)) + 8u * j);
            const uint32_t a_g = a_tab + uint32_t(offsetof(Q2Tab, G)) + cls * 256u + j * 4u;
            float am[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) am[k] = fmaxf(fabsf(lo2(w[k])), fabsf(hi2(w[k])));
            if (j == 0) am[0] = 0.0f;                                   // the DC coefficient went its own way
            const float hi_rows = fmaxf(fmaxf(fmaxf(am[3], am[4]), fmaxf(am[5], am[6])), am[7]);
            const bool any_hi = __reduce_max_sync(0xffffffffu, valid ? __float_as_uint(hi_rows) : 0u) >= thr;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k >= 3 && !any_hi) break;
                if (__reduce_max_sync(0xffffffffu, valid ? __float_as_uint(am[k]) : 0u) < thr) continue;
                const uint32_t zz = __byte_perm(k < 4 ? izz.x : izz.y, 0, 0x4440 + (k & 3));
                const float g = __uint_as_float(lds32(a_g + 32u * k));
                const f32x2 mg = pk2(kMagic15, kMagic15);
                const f32x2 dl = sub2(w[k], sub2(add2(w[k], mg), mg));          // w - rint(w)
                const uint32_t a_o = a_outA + 2u * zz, gbit = 1u << (zz >> 3);
                if ((k | j) != 0 && valid) {
                    const float wa = lo2(w[k]), wb = hi2(w[k]);
                    const int qa = __float2int_rz(wa), qb = __float2int_rz(wb);
                    if (fabsf(lo2(dl)) < g && fabsf(wa) > 0.5f) push_fix_w(a_cnt, a_fix, ((blkA) << 6) | uint32_t(k * 8) | j);
                    if (fabsf(hi2(dl)) < g && fabsf(wb) > 0.5f) push_fix_w(a_cnt, a_fix, ((blkA + (dAB >> 7)) << 6) | uint32_t(k * 8) | j);
                    if (qa != 0) sts16(a_o, qa), gm |= gbit;
                    if (qb != 0) sts16(a_o + dAB, qb), gm |= gbit << 8;
                }
            }
        }
        __syncwarp();

        // ---- phase 2b: this warp's guard-band queue, FP64, eight lanes per entry ----
        {
            const uint32_t nfix = lds32(a_cnt);
            if (nfix) {
                const uint32_t s8 = uint32_t(lane) & 7u;
                const bool overflow = nfix > uint32_t(kWarpFix);
                // overflow (adversarial content): every AC coefficient of the warp's eight blocks
                const uint32_t nent = overflow ? 8u * 63u : nfix;
                const uint32_t wblk0 = luma ? role * 12u : (role - 2u * T / 4u) * 24u + 4u;      // first block of the warp's pairs
                for (uint32_t e0 = 0; e0 < nent; e0 += 4) {
                    const uint32_t e = e0 + (uint32_t(lane) >> 3);
                    bool actv = e < nent;
                    uint32_t ent;
                    if (overflow) {
                        const uint32_t b8 = e / 63u, ij = e - b8 * 63u + 1u;     // the warp's b8-th block: luma m*6 + {0,1,2,3} of two MCUs, chroma m*6 + {4,5} of four
                        const uint32_t blk = luma ? wblk0 + (b8 >> 2) * 6u + (b8 & 3u) : wblk0 + (b8 >> 1) * 6u + (b8 & 1u);
                        ent = (blk << 6) | ij;
                    } else {
                        ent = actv ? lds32(a_fix + 2u * (e & ~1u)) >> (16u * (e & 1u)) & 0xffffu : 0u;
                    }
                    const uint32_t blk = (ent >> 6) & 0x7fu, ij = ent & 63u, ci = ij >> 3, cj = ij & 7u;
                    actv = actv && blk / 6u < nvalid;
                    const double* cjp = &gCosRef[cj * 8];
                    double row = 0.0;
#pragma unroll 1
                    for (int x = 0; x < 8; ++x) row = fma(double(tile_sample_s<C::kRow>(a_in, actv ? blk : blkA, int(s8), x, hlast, p.gray)), cjp[x], row);
                    double part = row * gCosRef[ci * 8 + s8];
                    part += __shfl_xor_sync(0xffffffffu, part, 1);
                    part += __shfl_xor_sync(0xffffffffu, part, 2);
                    part += __shfl_xor_sync(0xffffffffu, part, 4);
                    if (actv && s8 == 0) {
                        const int q = cC.quant[(blk % 6u) >= 4u][ij];
                        const int v = requant_finish_s<C::kRow>(part, a_in, blk, int(ci), int(cj), q, hlast, p.gray, p.guard_counter);
                        const uint32_t zz = cC.izz[ij];
                        uint32_t a_b = sm + C::oOut + ob * C::kOut + blk * 128u;
                        if constexpr (WST) {
                            const uint32_t bm = blk / 6u, bk = blk - bm * 6u;        // the warp's own blocks only
                            a_b = a_wst + (luma ? (bm - role * 2u) * 512u + bk * 128u : (bm - (role - 2u * T / 4u) * 4u) * 256u + (bk - 4u) * 128u);
                        }
                        sts16(a_b + 2u * zz, v);
                    }
                }
                __syncwarp();
                // a fixed-up coefficient may have become non-zero: the side information is recomputed from the staging buffer
                {
                    const uint32_t a_o = a_outA + sub * 16u;
                    const uint4 ga = lds128(a_o), gb = lds128(a_o + dAB);
                    const bool nza = ((sub ? ga.x : (ga.x >> 16)) | ga.y | ga.z | ga.w) != 0u, nzb = ((sub ? gb.x : (gb.x >> 16)) | gb.y | gb.z | gb.w) != 0u;
                    const uint32_t ba = __ballot_sync(0xffffffffu, nza), bb = __ballot_sync(0xffffffffu, nzb);
                    gm = ((ba >> (lane & 24)) & 0xffu) | (((bb >> (lane & 24)) & 0xffu) << 8);
                }
                if (lane == 0) sts32(a_cnt, 0u);
            } else if (__any_sync(0xffffffffu, gm != 0u)) {
                gm |= __shfl_xor_sync(0xffffffffu, gm, 1);
                gm |= __shfl_xor_sync(0xffffffffu, gm, 2);
                gm |= __shfl_xor_sync(0xffffffffu, gm, 4);
            }
        }
        // side information of the entropy coder: DC coefficient + non-zero groups of both blocks, straight from the registers
        if (p.bmeta && sub == 0 && valid) {
            uint32_t* meta = p.bmeta + size_t(img) * (p.coef_stride >> 6) + (size_t(my) * p.HU + mx0) * 6 + blkA;
            meta[0] = (uint32_t(dcA) & 0xffffu) | ((gm & 0xffu) << 16);
            meta[dAB >> 7] = (uint32_t(dcB) & 0xffffu) | ((gm & 0xff00u) << 8);
        }
        // done with the stage and with the staging buffer: tell the DMA warp
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            if constexpr (WST) {
                int16_t* dst = p.coefs + size_t(img) * p.coef_stride + (size_t(my) * p.HU + mx0) * 384;
                if (luma) {
#pragma unroll
                    for (uint32_t m = 0; m < 2; ++m)
                        if (role * 2u + m < nvalid) bulk_s2g(dst + (role * 2u + m) * 384u, a_wst + m * 512u, 512u);
                } else {
                    const uint32_t m0 = (role - 2u * T / 4u) * 4u;
#pragma unroll
                    for (uint32_t m = 0; m < 4; ++m)
                        if (m0 + m < nvalid) bulk_s2g(dst + (m0 + m) * 384u + 256u, a_wst + m * 256u, 256u);
                }
                bulk_commit();
            } else {
                mbar_arrive(bar_out_full + 8u * ob);
            }
            mbar_arrive(bar_in_empty + 8u * st);
        }
        if (++st == NST) st = 0, stpar ^= 1u;
        role += rot;
        if (role >= uint32_t(C::kWarps)) role -= uint32_t(C::kWarps);
        advance(img, my, bx);
    }
    if constexpr (WST) {
        if (lane == 0) bulk_wait_read0();        // shared memory must outlive the reads of the last stores
    }
}

}  // namespace jz
