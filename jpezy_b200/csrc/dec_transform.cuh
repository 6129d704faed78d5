// dec_transform.cuh -- stages D2+D3: quantised coefficients -> planar RGB.
//
// Replaces inverse_quantization (src/decoder/jpezy_decoder.hpp:645-650), inverse_dct (:652-670), the
// pixel replication of decode_mcu (:519-524) and make_rgb / to_r,g,b / revise_value (:531-578, :672-676).
//
// Numerics: the reference truncates int(sum/4 + 128) of an FP64 sum and then truncates the FP64 colour
// formulas, so results flip on 1-ulp differences when an exact value is an integer -- which is the common
// case for DC-only blocks and for neutral chroma.  The production kernel computes the IDCT in FP32 (AAN
// flowgraph) and decides every sample whose value is further than a guard band from an integer; the guard
// band is the worst-case FP32 error of the flowgraph for that block (sum of |coefficient| * sensitivity,
// tools/aan_idct_error_bound.py).  Samples inside the band, and whole DC-only blocks, go to a shared-memory
// queue and are re-evaluated densely in FP64, in the reference's exact operation order when they are within
// 1e-9 of an integer.  The colour conversion is FP32 with the same kind of argument (the exact values are
// multiples of 1e-4; R and B can only be integers when the chroma term is exactly 0), falling back to the
// reference's FP64 expression for the few chroma pairs whose G term is within 7.5e-5 of an integer.
// Result: decoded samples identical to the reference decoder's.  Algorithmic HBM traffic: 3 B/px read
// (int16 coefficients) + 3 B/px written = 6 B/px.
#pragma once
#include "common.cuh"
#include "enc_transform.cuh"

namespace jz {

struct InvParams {
    const int16_t* coefs;
    const int16_t* dc;        // != nullptr: [nimg][nblk] DC coefficients kept apart from `coefs` (whose position 0 is then 0)
    size_t coef_stride;
    uint8_t *r, *g, *b;
    size_t plane_stride;      // bytes between images (= plane_bytes)
    uint32_t W, H, HU, VU;    // HU x VU MCUs of (8*hmax) x (8*vmax) pixels
    int gray;
    // general frame layout (k_inv_transform_f64): blocks per MCU, luma blocks per MCU, sampling factors
    uint32_t row0;            // first MCU row of this launch (MCU-row shards of one image)
    uint32_t nb, ny, ncomp, hmax, vmax;
    uint32_t hs[3], vs[3];
    unsigned long long* guard_counter;
    uint16_t qt[3][64];       // per component, natural order
    float M[3][64];           // q * aan_v * aan_u / 8: dequantisation folded with the AAN input scaling
    float Wg[3][64];          // q * E * 2^-24 * 1.25: guard-band contribution per unit |coefficient|
};

constexpr double kGuardInv = 1e-6;

// exact-order IDCT sample (src/decoder/jpezy_decoder.hpp:657-668):
//   sum += cu*cv*dct[v*8+u]*cos[u*8+x]*cos[v*8+y]   (v outer, u inner, left-to-right products)
//   block[y*8+x] = int(sum / 4 + 128)
__device__ __noinline__ int idct_exact(const int* __restrict__ dq, unsigned long long nzmask, int x, int y)
{
    double sum = 0.0;
    while (nzmask) {
        const int pos = __ffsll((long long)nzmask) - 1;   // natural order == v*8+u ascending == reference order
        nzmask &= nzmask - 1;
        const int v = pos >> 3, u = pos & 7;
        const double cu = u ? 1.0 : cC.inv_sqrt2_ref, cv = v ? 1.0 : cC.inv_sqrt2_ref;
        double t = __dmul_rn(cu, cv);
        t = __dmul_rn(t, double(dq[pos]));
        t = __dmul_rn(t, cC.cos_ref[u * 8 + x]);
        t = __dmul_rn(t, cC.cos_ref[v * 8 + y]);
        sum = __dadd_rn(sum, t);
    }
    return __double2int_rz(__dadd_rn(__dmul_rn(sum, 0.25), 128.0));
}

// colour conversion, bit-exact with to_r/to_g/to_b + revise_value (:567-578, :672-676)
__device__ __forceinline__ uint8_t revise(double v) { return v < 0.0 ? uint8_t(0) : (v > 255.0 ? uint8_t(255) : uint8_t(__double2int_rz(v))); }
__device__ __forceinline__ uint8_t ref_R(int y, int cr) { return revise(__dadd_rn(double(y), __dmul_rn(double(cr - 128), 1.4020))); }
__device__ __forceinline__ uint8_t ref_G(int y, int cb, int cr)
{
    const double t = __dsub_rn(double(y), __dmul_rn(double(cb - 128), 0.3441));
    return revise(__dsub_rn(t, __dmul_rn(double(cr - 128), 0.7139)));
}
__device__ __forceinline__ uint8_t ref_B(int y, int cb) { return revise(__dadd_rn(double(y), __dmul_rn(double(cb - 128), 1.7718))); }

// ---- general layouts + validation build: FP64 separable IDCT -----------------------------------------------
// Handles every frame layout the device decoder accepts (1 or 3 components, luma H, V in {1, 2}, chroma 1x1): the
// production kernel below is specialised for jpezy's own 2x2 / 1x1 / 1x1.  One CTA = kMcuPerCta MCUs of one MCU row.
// Pixel replication of decode_mcu (src/decoder/jpezy_decoder.hpp:519-524): sample (py / dupy, px / dupx) of the
// component's blocks, dup = max factor / component factor.
__global__ void __launch_bounds__(kFwdThreads) k_inv_transform_f64(const InvParams p)
{
    pdl_wait();
    __shared__ int s_dq[kBlkPerCta][64];                 // dequantised, natural order
    __shared__ double s_tmp[kBlkPerCta][64];
    __shared__ short s_pix[kBlkPerCta][64];              // IDCT output (unclamped int, fits 16 bits)
    __shared__ unsigned long long s_nz[kBlkPerCta];

    const int t = threadIdx.x;
    const uint32_t mx0 = blockIdx.x * kMcuPerCta;
    const uint32_t my = blockIdx.y + p.row0;
    const size_t img = blockIdx.z;
    const uint32_t nvalid = min(uint32_t(kMcuPerCta), p.HU - mx0);
    const uint32_t nb = p.nb, ny = p.ny, nblk = kMcuPerCta * nb;      // nb <= 6

    const int16_t* src = p.coefs + img * p.coef_stride + (size_t(my) * p.HU + mx0) * nb * 64;
    for (uint32_t e = t; e < nblk * 64; e += kFwdThreads) {
        const uint32_t blk = e >> 6, n = e & 63;
        const uint32_t k = blk % nb;
        const int comp = k < ny ? 0 : int(k - ny) + 1;
        const int nat = cC.zz[n];
        const int c = (blk / nb) < nvalid ? int(src[e]) : 0;
        s_dq[blk][nat] = c * int(p.qt[comp][nat]);
    }
    __syncthreads();
    if (uint32_t(t) < nblk) {
        unsigned long long m = 0;
        for (int i = 0; i < 64; ++i)
            if (s_dq[t][i]) m |= 1ull << i;
        s_nz[t] = m;
    }
    for (uint32_t task = t; task < nblk * 8; task += kFwdThreads) {
        const int blk = task >> 3, u = task & 7;
        double col[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) col[v] = double(s_dq[blk][v * 8 + u]) * (v ? 1.0 : 0.70710678118654752440);
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            double s = 0.0;
#pragma unroll
            for (int v = 0; v < 8; ++v) s = fma(col[v], cC.cos_ref[v * 8 + y], s);
            s_tmp[blk][y * 8 + u] = s;
        }
    }
    __syncthreads();
    unsigned long long guard_hits = 0;
    for (uint32_t task = t; task < nblk * 8; task += kFwdThreads) {
        const int blk = task >> 3, y = task & 7;
        double row[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) row[u] = s_tmp[blk][y * 8 + u] * (u ? 1.0 : 0.70710678118654752440);
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) s = fma(row[u], cC.cos_ref[u * 8 + x], s);
            const double val = s * 0.25 + 128.0;
            int iv;
            if (fabs(val - rint(val)) < kGuardInv) {
                iv = idct_exact(&s_dq[blk][0], s_nz[blk], x, y);
                ++guard_hits;
            } else {
                iv = __double2int_rz(val);
            }
            s_pix[blk][y * 8 + x] = short(iv);
        }
    }
    __syncthreads();
    if (guard_hits) atomicAdd(p.guard_counter, guard_hits);

    uint8_t* R = p.r + img * p.plane_stride;
    uint8_t* G = p.g + img * p.plane_stride;
    uint8_t* B = p.b + img * p.plane_stride;
    const uint32_t mw = 8 * p.hmax, mh = 8 * p.vmax;             // MCU size in pixels
    const uint32_t dx1 = p.ncomp == 3 ? p.hmax / p.hs[1] : 1, dy1 = p.ncomp == 3 ? p.vmax / p.vs[1] : 1;
    const uint32_t dx2 = p.ncomp == 3 ? p.hmax / p.hs[2] : 1, dy2 = p.ncomp == 3 ? p.vmax / p.vs[2] : 1;
    const uint32_t dx0 = p.hmax / p.hs[0], dy0 = p.vmax / p.vs[0];
    for (uint32_t e = t; e < mh * kMcuPerCta * mw; e += kFwdThreads) {
        const uint32_t ry = e / (kMcuPerCta * mw), cx = e - ry * (kMcuPerCta * mw);
        const uint32_t mcu = cx / mw, px = cx - mcu * mw;
        const uint32_t gx = (mx0 + mcu) * mw + px;
        if (mcu >= nvalid || gx >= p.W) continue;
        const uint32_t sx0 = px / dx0, sy0 = ry / dy0;
        const int yv = s_pix[mcu * nb + (sy0 >> 3) * p.hs[0] + (sx0 >> 3)][(sy0 & 7) * 8 + (sx0 & 7)];
        const size_t idx = (size_t(my) * mh + ry) * p.W + gx;
        if (p.gray) {
            const uint8_t v = revise(double(yv));
            R[idx] = v, G[idx] = v, B[idx] = v;
        } else {
            int cb = 128, cr = 128;      // comp tiles of absent components stay 0x80 (src/decoder/jpezy_decoder.hpp:105)
            if (p.ncomp == 3) {
                const uint32_t sx1 = px / dx1, sy1 = ry / dy1, sx2 = px / dx2, sy2 = ry / dy2;
                cb = s_pix[mcu * nb + ny + (sy1 >> 3) * p.hs[1] + (sx1 >> 3)][(sy1 & 7) * 8 + (sx1 & 7)];
                cr = s_pix[mcu * nb + ny + p.hs[1] * p.vs[1] + (sy2 >> 3) * p.hs[2] + (sx2 >> 3)][(sy2 & 7) * 8 + (sx2 & 7)];
            }
            R[idx] = ref_R(yv, cr), G[idx] = ref_G(yv, cb, cr), B[idx] = ref_B(yv, cb);
        }
    }
}

// =====================================================================================================
// Production kernel
//
//  * tile = 16 rows x 512 pixels (32 MCUs, 192 blocks) per CTA of 192 threads, 3 CTAs/SM
//  * phase 0: 16-byte coalesced loads of the tile's coefficients into a padded staging buffer; a ballot over the 8
//    chunks of a block yields the mask of its non-zero zig-zag groups
//  * phase 1: one thread per block (warps 0..3 luma, 4 Cb, 5 Cr): dequantisation folded with the AAN input scaling, FP32
//    AAN IDCT in registers.  When no block of the warp has a coefficient beyond zig-zag position 7 the transform is
//    pruned (4 three-input column passes + 8 four-input row passes, operations on structural zeros dropped, so the
//    results are bit-identical to the full flowgraph).  A sample is decided when it is further than the block's guard
//    band from an integer (one min-reduction per row of 8); the others go to the fix-up queue.  DC-only blocks are
//    evaluated in the reference's operation order right away.
//  * phase 1b: fix-up queue, FP64 (exact order when within 1e-9 of an integer)
//  * phase 1c: the 2048 chroma pairs of the tile -> integer offsets floor((Cr-128)*1.4020), floor(-(Cb-128)*0.3441 -
//    (Cr-128)*0.7139), floor((Cb-128)*1.7718): trunc(y + t) = y + floor(t) for every value that is not clamped, and the
//    clamp absorbs the rest (tools/colour_floor_check.py proves this exhaustively); pairs whose G term is within 1.2e-4
//    of an integer are flagged and take the reference's FP64 expression per pixel
//  * phase 2: one thread per (row, MCU): 16 pixels, integer adds, saturating packs, three 16-byte planar stores
// =====================================================================================================
constexpr int kInvThreads = 192;
constexpr int kIYStride = 1040;    // bytes per row of the luma sample tile (512 int16 + 16)
constexpr int kICStride = 528;     // bytes per row of a chroma sample tile (256 int16 + 16)
constexpr int kInvSmem = kTileBlk * kOutStride + 16 * kIYStride + 2 * 8 * kICStride + kFixCap * 2 + 16 + kTileBlk;
static_assert(2 * 2048 * 4 <= kTileBlk * kOutStride, "the chroma offset tables reuse the coefficient staging buffer");

// worst-case output error of the FP32 AAN inverse flowgraph per unit of dequantised coefficient, in units of
// 2^-24 (tools/aan_idct_error_bound.py), natural order, rounded up
static const float kIdctErrSens[64] = {
    1.0f, 11.5f, 3.1f, 11.7f, 1.1f, 8.1f, 1.7f, 2.9f,
    7.9f, 78.4f, 24.7f, 80.6f, 7.9f, 54.0f, 10.4f, 15.8f,
    2.1f, 21.3f, 6.5f, 21.8f, 2.1f, 14.7f, 2.9f, 4.6f,
    8.1f, 82.0f, 25.5f, 84.2f, 8.1f, 56.5f, 10.9f, 16.8f,
    1.1f, 11.5f, 3.1f, 11.7f, 1.1f, 8.1f, 1.7f, 2.9f,
    5.9f, 60.5f, 18.5f, 62.1f, 5.9f, 41.8f, 8.2f, 12.7f,
    1.7f, 18.6f, 5.1f, 18.9f, 1.7f, 13.0f, 2.7f, 4.4f,
    2.6f, 27.9f, 8.0f, 28.5f, 2.6f, 19.4f, 3.9f, 6.3f};

// AAN inverse 1-D DCT on 8 registers (inputs pre-scaled by aan_k, see InvParams::M)
__device__ __forceinline__ void aan_idct8(float& d0, float& d1, float& d2, float& d3, float& d4, float& d5, float& d6, float& d7)
{
    const float t10 = d0 + d4, t11 = d0 - d4;
    const float t13 = d2 + d6;
    const float t12 = fmaf(d2 - d6, 1.414213562373095049f, -t13);
    const float e0 = t10 + t13, e3 = t10 - t13, e1 = t11 + t12, e2 = t11 - t12;
    const float z13 = d5 + d3, z10 = d5 - d3, z11 = d1 + d7, z12 = d1 - d7;
    const float o7 = z11 + z13;
    const float o11 = (z11 - z13) * 1.414213562373095049f;
    const float z5 = (z10 + z12) * 1.847759065022573512f;
    const float o10 = fmaf(z12, 1.082392200292393968f, -z5);
    const float o12 = fmaf(z10, -2.613125929752753055f, z5);
    const float o6 = o12 - o7;
    const float o5 = o11 - o6;
    const float o4 = o10 + o5;
    d0 = e0 + o7, d7 = e0 - o7;
    d1 = e1 + o6, d6 = e1 - o6;
    d2 = e2 + o5, d5 = e2 - o5;
    d4 = e3 + o4, d3 = e3 - o4;
}
// the same flowgraph with d3..d7 == 0 (resp. d4..d7 == 0): every operation whose operand is a structural zero is dropped
// (x + 0, x - 0, 0 * c are exact), so the results equal aan_idct8's bit for bit
__device__ __forceinline__ void aan_idct8_in3(float& d0, float& d1, float& d2, float& d3, float& d4, float& d5, float& d6, float& d7)
{
    const float t12 = fmaf(d2, 1.414213562373095049f, -d2);
    const float e0 = d0 + d2, e3 = d0 - d2, e1 = d0 + t12, e2 = d0 - t12;
    const float o7 = d1;
    const float o11 = d1 * 1.414213562373095049f;
    const float z5 = d1 * 1.847759065022573512f;
    const float o10 = fmaf(d1, 1.082392200292393968f, -z5);
    const float o6 = z5 - o7;
    const float o5 = o11 - o6;
    const float o4 = o10 + o5;
    d0 = e0 + o7, d7 = e0 - o7;
    d1 = e1 + o6, d6 = e1 - o6;
    d2 = e2 + o5, d5 = e2 - o5;
    d4 = e3 + o4, d3 = e3 - o4;
}
__device__ __forceinline__ void aan_idct8_in4(float& d0, float& d1, float& d2, float& d3, float& d4, float& d5, float& d6, float& d7)
{
    const float t12 = fmaf(d2, 1.414213562373095049f, -d2);
    const float e0 = d0 + d2, e3 = d0 - d2, e1 = d0 + t12, e2 = d0 - t12;
    const float dm = d1 - d3;                                   // z11 - z13 == z10 + z12 (z13 = d3, z10 = -d3, z11 = z12 = d1)
    const float o7 = d1 + d3;
    const float o11 = dm * 1.414213562373095049f;
    const float z5 = dm * 1.847759065022573512f;
    const float o10 = fmaf(d1, 1.082392200292393968f, -z5);
    const float o12 = fmaf(d3, 2.613125929752753055f, z5);       // (-d3) * (-2.613...) + z5
    const float o6 = o12 - o7;
    const float o5 = o11 - o6;
    const float o4 = o10 + o5;
    d0 = e0 + o7, d7 = e0 - o7;
    d1 = e1 + o6, d6 = e1 - o6;
    d2 = e2 + o5, d5 = e2 - o5;
    d4 = e3 + o4, d3 = e3 - o4;
}

// dequantise one 16-byte group (8 zig-zag consecutive coefficients) into the natural-order register array
template <int GRP>
__device__ __forceinline__ void dequant_group(const float* __restrict__ M, const float* __restrict__ Wg, const uint4 raw, float (&d)[64], float& gsum)
{
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int h = 0; h < 8; ++h) {
        const int nat = zz_at(GRP * 8 + h);
        const int c = (h & 1) ? (int(w[h >> 1]) >> 16) : int(short(w[h >> 1] & 0xffffu));
        const float cf = float(c);
        d[nat] = cf * M[nat];
        gsum = fmaf(fabsf(cf), Wg[nat], gsum);
    }
}

// dequantisation + IDCT of one block; `wm` = OR of the non-zero-group masks of the warp's blocks (warp uniform)
// (comp is warp uniform: the tables are read from the constant bank with a uniform offset; one copy of the code for the
// three components keeps the kernel's instruction footprint down)
__device__ __forceinline__ void idct_block(const InvParams& p, const int comp, const uint4* __restrict__ src, const uint4 raw0, const uint32_t wm,
                                           float (&d)[64], float& gsum)
{
    const float* __restrict__ M = p.M[comp];
    const float* __restrict__ Wg = p.Wg[comp];
    gsum = 2e-5f;
    if (wm <= 1u) {
        // only zig-zag positions 0..7 = natural (0,0) (0,1) (1,0) (2,0) (1,1) (0,2) (0,3) (1,2): rows v <= 2, columns u <= 3
        dequant_group<0>(M, Wg, raw0, d, gsum);
        d[11] = d[17] = d[18] = d[19] = 0.0f;
        d[0] += 128.0f;   // level shift rides on the DC term (gain 1 through the flowgraph)
#pragma unroll
        for (int u = 0; u < 4; ++u) aan_idct8_in3(d[u], d[8 + u], d[16 + u], d[24 + u], d[32 + u], d[40 + u], d[48 + u], d[56 + u]);
#pragma unroll
        for (int y = 0; y < 8; ++y)
            aan_idct8_in4(d[y * 8 + 0], d[y * 8 + 1], d[y * 8 + 2], d[y * 8 + 3], d[y * 8 + 4], d[y * 8 + 5], d[y * 8 + 6], d[y * 8 + 7]);
        return;
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) d[k] = 0.0f;
    dequant_group<0>(M, Wg, raw0, d, gsum);
    if (wm & 0x02u) dequant_group<1>(M, Wg, src[1], d, gsum);
    if (wm & 0x04u) dequant_group<2>(M, Wg, src[2], d, gsum);
    if (wm & 0x08u) dequant_group<3>(M, Wg, src[3], d, gsum);
    if (wm & 0x10u) dequant_group<4>(M, Wg, src[4], d, gsum);
    if (wm & 0x20u) dequant_group<5>(M, Wg, src[5], d, gsum);
    if (wm & 0x40u) dequant_group<6>(M, Wg, src[6], d, gsum);
    if (wm & 0x80u) dequant_group<7>(M, Wg, src[7], d, gsum);
    d[0] += 128.0f;
#pragma unroll
    for (int u = 0; u < 8; ++u) aan_idct8(d[u], d[8 + u], d[16 + u], d[24 + u], d[32 + u], d[40 + u], d[48 + u], d[56 + u]);
#pragma unroll
    for (int y = 0; y < 8; ++y)
        aan_idct8(d[y * 8 + 0], d[y * 8 + 1], d[y * 8 + 2], d[y * 8 + 3], d[y * 8 + 4], d[y * 8 + 5], d[y * 8 + 6], d[y * 8 + 7]);
}

// FP64 re-evaluation of one sample of block `cz` (zig-zag int16 coefficients, non-zero only below position nlim) with
// quantiser table qt (natural order)
__device__ __noinline__ int idct_fix_finish(double acc, const int16_t* __restrict__ cz, const uint16_t* __restrict__ qt, int nlim, int x, int y,
                                            uint32_t* exact_hits)
{
    const double val = acc * 0.25 + 128.0;
    if (fabs(val - rint(val)) >= 1e-9) return __double2int_rz(val);
    // the reference's exact operation order: natural order ascending, zero terms cannot change the sum
    unsigned long long nzmask = 0;
    for (int n = 0; n < nlim; ++n)
        if (cz[n]) nzmask |= 1ull << cC.zz[n];
    double sum = 0.0;
    while (nzmask) {
        const int nat = __ffsll((long long)nzmask) - 1;
        nzmask &= nzmask - 1;
        const int v = nat >> 3, u = nat & 7;
        const double cu = u ? 1.0 : cC.inv_sqrt2_ref, cv = v ? 1.0 : cC.inv_sqrt2_ref;
        double t = __dmul_rn(cu, cv);
        t = __dmul_rn(t, double(int(cz[cC.izz[nat]]) * int(qt[nat])));
        t = __dmul_rn(t, cC.cos_ref[u * 8 + x]);
        t = __dmul_rn(t, cC.cos_ref[v * 8 + y]);
        sum = __dadd_rn(sum, t);
    }
    ++*exact_hits;
    return __double2int_rz(__dadd_rn(__dmul_rn(sum, 0.25), 128.0));
}

// partial FP64 sum of sample (x, y) over the zig-zag positions first, first + step, ... below nlim
__device__ __forceinline__ double idct_fix_part(const int16_t* __restrict__ cz, const uint16_t* __restrict__ qt, int first, int step, int nlim, int x, int y)
{
    double acc = 0.0;
    for (int n = first; n < nlim; n += step) {
        const int c = cz[n];
        if (!c) continue;
        const int nat = cC.zz[n], v = nat >> 3, u = nat & 7;
        const double f = double(c * int(qt[nat])) * (u ? 1.0 : 0.70710678118654752440) * (v ? 1.0 : 0.70710678118654752440);
        acc = fma(f * gCosRef[u * 8 + x], gCosRef[v * 8 + y], acc);
    }
    return acc;
}

__device__ __noinline__ int idct_fix(const int16_t* __restrict__ cz, const uint16_t* __restrict__ qt, int nlim, int x, int y, uint32_t* exact_hits)
{
    return idct_fix_finish(idct_fix_part(cz, qt, 0, 1, nlim, x, y), cz, qt, nlim, x, y, exact_hits);
}

// colour conversion of 4 horizontally adjacent pixels sharing 2 chroma pairs; y4: 4 int16 in two words
__device__ __noinline__ uint32_t colour_exact4(const int* y, int cb0, int cr0, int cb1, int cr1, int which)
{
    uint32_t out = 0;
    for (int i = 0; i < 4; ++i) {
        const int cb = i < 2 ? cb0 : cb1, cr = i < 2 ? cr0 : cr1;
        const uint32_t v = which == 0 ? ref_R(y[i], cr) : (which == 1 ? ref_G(y[i], cb, cr) : ref_B(y[i], cb));
        out |= v << (8 * i);
    }
    return out;
}

constexpr uint32_t kWholeBlock = 1u << 15;   // above the 8-bit block number (blk << 7 occupies bits 7..14)

__device__ __noinline__ void push_fix16(uint32_t* s_nfix, uint16_t* s_fix, uint32_t entry)
{
    const uint32_t idx = atomicAdd(s_nfix, 1u);
    if (idx < kFixCap) s_fix[idx] = uint16_t(entry);
}

// slow half of the guard test of one row: which of the 8 samples are inside the band
__device__ __noinline__ void flag_row(const float (&v)[8], float guard, uint32_t blk, int y, uint32_t* s_nfix, uint16_t* s_fix)
{
#pragma unroll 1
    for (int x = 0; x < 8; ++x) {
        const float kf = (v[x] + 12582912.0f) - 12582912.0f;
        if (fabsf(v[x] - kf) < guard) push_fix16(s_nfix, s_fix, (blk << 7) | uint32_t(y * 8 + x));
    }
}

// four s32 -> four saturated u8 in one word (v0 in the low byte): d = (c[15:0] << 16) | sat(a) << 8 | sat(b)
__device__ __forceinline__ uint32_t pack_sat_u8x4(int v0, int v1, int v2, int v3)
{
    uint32_t hi, out;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(v3), "r"(v2));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(v1), "r"(v0), "r"(hi));
    return out;
}

__device__ __forceinline__ int sx_lo(uint32_t w) { return int(short(w & 0xffffu)); }   // sign-extended low half (one SGXT / PRMT)
__device__ __forceinline__ int sx_hi(uint32_t w) { return int(w) >> 16; }

__global__ void __launch_bounds__(kInvThreads, 3) k_inv_transform(const __grid_constant__ InvParams p)
{
    pdl_wait();
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_coef = smem;                                  // [192][144] zig-zag int16 coefficients (padded); later the offset tables
    uint8_t* s_y = s_coef + kTileBlk * kOutStride;           // [16][kIYStride] int16 luma samples
    uint8_t* s_cb = s_y + 16 * kIYStride;                    // [8][kICStride]
    uint8_t* s_cr = s_cb + 8 * kICStride;
    uint16_t* s_fix = reinterpret_cast<uint16_t*>(s_cr + 8 * kICStride);
    uint32_t* s_nfix = reinterpret_cast<uint32_t*>(s_fix + kFixCap);
    uint8_t* s_mask = reinterpret_cast<uint8_t*>(s_nfix) + 16;   // [192] non-zero zig-zag groups of every block

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t mx0 = blockIdx.x * kTileMcu;
    const uint32_t my = blockIdx.y + p.row0;
    const size_t img = blockIdx.z;
    const uint32_t nvalid = min(uint32_t(kTileMcu), p.HU - mx0);
    if (t == 0) *s_nfix = 0;
    // the block this thread transforms in phase 1 (warps 0..3 luma, 4 Cb, 5 Cr); luma: warp 0 = top blocks of MCUs 0..15,
    // warp 1 = their bottom blocks, warps 2/3 = MCUs 16..31
    const uint32_t my_blk = warp < 4 ? ((warp >> 1) * 16 + (lane >> 1)) * 6 + (warp & 1) * 2 + (lane & 1) : uint32_t(lane) * 6 + warp;
    // its DC coefficient when the entropy decoder kept those in the dense side array (DecParams::dcd): position 0 of the
    // block is 0 in `coefs` then; fetched now, merged in phase 1
    uint32_t my_dc = 0;
    if (p.dc && my_blk < nvalid * 6u) my_dc = uint16_t(__ldg(p.dc + img * (p.coef_stride >> 6) + (size_t(my) * p.HU + mx0) * 6 + my_blk));

    // ---- phase 0: coalesced load of the tile's coefficients into the padded staging buffer ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.coefs + img * p.coef_stride + (size_t(my) * p.HU + mx0) * 384);
        const uint32_t nchunks = nvalid * 48;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t c = uint32_t(t) + uint32_t(i) * kInvThreads;
            const uint4 v = c < nchunks ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(&s_coef[(c >> 3) * kOutStride + (c & 7) * 16]) = v;
            const uint32_t nzb = __ballot_sync(0xffffffffu, (v.x | v.y | v.z | v.w) != 0u);
            if ((lane & 7) == 0) s_mask[c >> 3] = uint8_t(nzb >> lane);
        }
    }
    __syncthreads();

    // ---- phase 1: dequantisation + IDCT, one thread per block (warps 0..3 luma, 4 Cb, 5 Cr) ----
    {
        uint32_t blk;
        uint8_t* tile;
        int stride, comp;
        if (warp < 4) {
            const uint32_t mcu = (warp >> 1) * 16 + (lane >> 1), k = (warp & 1) * 2 + (lane & 1);
            blk = mcu * 6 + k;
            tile = &s_y[(warp & 1) * 8 * kIYStride + (mcu * 16 + (lane & 1) * 8) * 2];
            stride = kIYStride, comp = 0;
        } else {
            blk = lane * 6 + warp;
            tile = (warp == 4 ? s_cb : s_cr) + lane * 16;
            stride = kICStride, comp = warp - 3;
        }
        const uint32_t mask = s_mask[blk];
        const uint32_t wm = __reduce_or_sync(0xffffffffu, mask);
        const uint4* src = reinterpret_cast<const uint4*>(&s_coef[blk * kOutStride]);
        uint4 raw0 = src[0];
        if (p.dc) {                // (the fix-up phase reads the block from shared memory again)
            raw0.x |= my_dc;
            *reinterpret_cast<uint32_t*>(&s_coef[blk * kOutStride]) = raw0.x;
        }
        const bool dc_only = mask <= 1u && ((raw0.x >> 16) | raw0.y | raw0.z | raw0.w) == 0u;
        if (__all_sync(0xffffffffu, dc_only)) {
            // every block of the warp is DC-only: ((c*c)*F)*1*1, /4, +128 exactly as the reference evaluates it
            const double f = double(int(short(raw0.x & 0xffffu)) * int(p.qt[comp][0]));
            const double term = __dmul_rn(__dmul_rn(cC.inv_sqrt2_ref, cC.inv_sqrt2_ref), f);
            const int v = __double2int_rz(__dadd_rn(__dmul_rn(term, 0.25), 128.0));
            const uint32_t vv = __byte_perm(uint32_t(v), uint32_t(v), 0x5410);
#pragma unroll
            for (int y = 0; y < 8; ++y) *reinterpret_cast<uint4*>(tile + y * stride) = make_uint4(vv, vv, vv, vv);
            if (lane == 0) atomicAdd(p.guard_counter, 32ull * 64ull);
        } else {
            float d[64];
            float gsum;
            uint32_t dc_exact = 0;
            idct_block(p, comp, src, raw0, wm, d, gsum);
            if (dc_only) {
                // the exact value is very often an integer: ((c*c)*F)*1*1, /4, +128 exactly as the reference evaluates it
                const double f = double(int(short(raw0.x & 0xffffu)) * int(p.qt[comp][0]));
                const double term = __dmul_rn(__dmul_rn(cC.inv_sqrt2_ref, cC.inv_sqrt2_ref), f);
                const int v = __double2int_rz(__dadd_rn(__dmul_rn(term, 0.25), 128.0));
                const uint32_t vv = __byte_perm(uint32_t(v), uint32_t(v), 0x5410);
#pragma unroll
                for (int y = 0; y < 8; ++y) *reinterpret_cast<uint4*>(tile + y * stride) = make_uint4(vv, vv, vv, vv);
                dc_exact = 64;
            } else {
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    int iv[8];
                    float vr[8];
                    float m = 1.0f;
#pragma unroll
                    for (int x = 0; x < 8; ++x) {
                        const float val = d[y * 8 + x];
                        vr[x] = val;
                        iv[x] = __float2int_rz(val);
                        const float kf = (val + 12582912.0f) - 12582912.0f;   // rint(val), |val| < 2^22
                        m = fminf(m, fabsf(val - kf));
                    }
                    if (m < gsum) flag_row(vr, gsum, blk, y, s_nfix, s_fix);
                    uint4 v;
                    v.x = __byte_perm(uint32_t(iv[0]), uint32_t(iv[1]), 0x5410), v.y = __byte_perm(uint32_t(iv[2]), uint32_t(iv[3]), 0x5410);
                    v.z = __byte_perm(uint32_t(iv[4]), uint32_t(iv[5]), 0x5410), v.w = __byte_perm(uint32_t(iv[6]), uint32_t(iv[7]), 0x5410);
                    *reinterpret_cast<uint4*>(tile + y * stride) = v;
                }
            }
            dc_exact = __reduce_add_sync(0xffffffffu, dc_exact);
            if (lane == 0 && dc_exact) atomicAdd(p.guard_counter, (unsigned long long)dc_exact);
        }
    }
    __syncthreads();

    // ---- phase 1b: dense re-evaluation of the queue (guard-band samples) ----
    {
        const uint32_t nfix = *s_nfix;
        if (nfix) {
            const bool overflow = nfix > kFixCap;
            uint32_t exact_hits = 0;    // samples decided by the reference's exact operation order (JPEZYB200_STAT_GUARD_INV)
            // eight lanes per queue entry; the loop bound is CTA-uniform so that every lane reaches the butterfly
            const uint32_t ntask = overflow ? kTileBlk * 8u : nfix * 8u;
            for (uint32_t task0 = 0; task0 < ntask; task0 += kInvThreads) {
                const uint32_t task = task0 + uint32_t(t);
                const bool act = task < ntask;
                // entry = blk << 7 | sample; kWholeBlock (overflow only) = every sample of the block
                const uint32_t e = !act ? 0u : (overflow ? (((task >> 3) << 7) | kWholeBlock) : uint32_t(s_fix[task >> 3]));
                const uint32_t blk = (e >> 7) & 255u, sub = task & 7u;
                const uint32_t mcu = blk / 6u, k = blk - mcu * 6u;
                const int comp = k < 4u ? 0 : int(k) - 3;
                const int16_t* cz = reinterpret_cast<const int16_t*>(&s_coef[blk * kOutStride]);
                const int nlim = 8 * (32 - __clz(uint32_t(s_mask[blk]) | 1u));
                int16_t* tile;
                int stride;
                if (k < 4u) {
                    tile = reinterpret_cast<int16_t*>(&s_y[(k >> 1) * 8 * kIYStride + (mcu * 16 + (k & 1) * 8) * 2]);
                    stride = kIYStride / 2;
                } else {
                    tile = reinterpret_cast<int16_t*>((k == 4u ? s_cb : s_cr) + mcu * 16);
                    stride = kICStride / 2;
                }
                if (overflow) {            // row `sub` of the block, every sample in FP64
                    if (act)
                        for (int x = 0; x < 8; ++x) tile[sub * stride + x] = int16_t(idct_fix(cz, p.qt[comp], nlim, x, int(sub), &exact_hits));
                } else {                   // lane `sub` sums the zig-zag positions sub, sub + 8, ...; a 3-step butterfly adds the lanes
                    const int s = int(e & 63u);
                    double part = act ? idct_fix_part(cz, p.qt[comp], int(sub), 8, nlim, s & 7, s >> 3) : 0.0;
                    part += __shfl_xor_sync(0xffffffffu, part, 1);
                    part += __shfl_xor_sync(0xffffffffu, part, 2);
                    part += __shfl_xor_sync(0xffffffffu, part, 4);
                    if (act && sub == 0)
                        tile[(s >> 3) * stride + (s & 7)] = int16_t(idct_fix_finish(part, cz, p.qt[comp], nlim, s & 7, s >> 3, &exact_hits));
                }
            }
            exact_hits = __reduce_add_sync(0xffffffffu, exact_hits);
            if (lane == 0 && exact_hits) atomicAdd(p.guard_counter, (unsigned long long)exact_hits);
        }
    }
    __syncthreads();

    // ---- phase 1c: chroma pairs -> integer colour offsets (the coefficient staging buffer is free now) ----
    uint32_t* s_offa = reinterpret_cast<uint32_t*>(s_coef);            // [8][256]  fr | fg << 16
    uint32_t* s_offb = s_offa + 2048;                                  // [8][256]  fb | suspect << 16
    if (!p.gray) {
#pragma unroll 1
        for (uint32_t pp = t; pp < 1024u; pp += kInvThreads) {     // two horizontally adjacent pairs per iteration
            const uint32_t crow = pp >> 7, cx2 = pp & 127u;
            const uint32_t cbw = *reinterpret_cast<const uint32_t*>(&s_cb[crow * kICStride + cx2 * 4]);
            const uint32_t crw = *reinterpret_cast<const uint32_t*>(&s_cr[crow * kICStride + cx2 * 4]);
            uint32_t oa[2], ob[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cb = h ? sx_hi(cbw) : sx_lo(cbw), cr = h ? sx_hi(crw) : sx_lo(crw);
                const float a = float(cb - 128), b = float(cr - 128);
                const int fr = __float2int_rd(b * 1.4020f);
                const int fb = __float2int_rd(a * 1.7718f);
                const float tg = fmaf(b, -0.7139f, a * -0.3441f);
                const int fg = __float2int_rd(tg);
                // G is an exact integer only when 3441a + 7139b = 0 (mod 10000); FP32 error of tg < 4.5e-5 for |a|,|b| <= 256
                const float kf = (tg + 12582912.0f) - 12582912.0f;
                const bool near_int = fabsf(tg - kf) < 7.5e-5f && ((cb ^ 128) | (cr ^ 128)) != 0;
                const bool wild = (uint32_t(cb + 128) | uint32_t(cr + 128)) > 512u;   // (conservative: OR of two values <= 512)
                oa[h] = __byte_perm(uint32_t(fr), uint32_t(fg), 0x5410);
                ob[h] = (uint32_t(fb) & 0xffffu) | ((near_int || wild) ? 0x10000u : 0u);
            }
            *reinterpret_cast<uint2*>(&s_offa[pp * 2]) = make_uint2(oa[0], oa[1]);
            *reinterpret_cast<uint2*>(&s_offb[pp * 2]) = make_uint2(ob[0], ob[1]);
        }
        __syncthreads();
    }

    // ---- phase 2: colour + planar stores; one thread per (row, MCU) ----
    uint8_t* R = p.r + img * p.plane_stride;
    uint8_t* G = p.g + img * p.plane_stride;
    uint8_t* B = p.b + img * p.plane_stride;
#pragma unroll 1
    for (uint32_t task = t; task < 512u; task += kInvThreads) {
        const uint32_t ry = task >> 5, mcu = task & 31u;
        const uint32_t x0 = (mx0 + mcu) * 16u;
        if (mx0 + mcu >= p.HU) continue;
        const uint4 ya = *reinterpret_cast<const uint4*>(&s_y[ry * kIYStride + mcu * 32]);
        const uint4 yb = *reinterpret_cast<const uint4*>(&s_y[ry * kIYStride + mcu * 32 + 16]);
        const uint32_t yw[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
        uint32_t ro[4], go[4], bo[4];
        if (p.gray) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
                ro[q4] = go[q4] = bo[q4] = pack_sat_u8x4(sx_lo(yw[q4 * 2]), sx_hi(yw[q4 * 2]), sx_lo(yw[q4 * 2 + 1]), sx_hi(yw[q4 * 2 + 1]));
        } else {
            const uint32_t obase = (ry >> 1) * 256u + mcu * 8u;
            const uint4 oa0 = *reinterpret_cast<const uint4*>(&s_offa[obase]), oa1 = *reinterpret_cast<const uint4*>(&s_offa[obase + 4]);
            const uint4 ob0 = *reinterpret_cast<const uint4*>(&s_offb[obase]), ob1 = *reinterpret_cast<const uint4*>(&s_offb[obase + 4]);
            const uint32_t oa[8] = {oa0.x, oa0.y, oa0.z, oa0.w, oa1.x, oa1.y, oa1.z, oa1.w};
            const uint32_t ob[8] = {ob0.x, ob0.y, ob0.z, ob0.w, ob1.x, ob1.y, ob1.z, ob1.w};
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {       // 4 pixels = 2 chroma pairs
                const int c0 = q4 * 2, c1 = q4 * 2 + 1;
                const int y0 = sx_lo(yw[c0]), y1 = sx_hi(yw[c0]), y2 = sx_lo(yw[c1]), y3 = sx_hi(yw[c1]);
                if ((ob[c0] | ob[c1]) & 0x10000u) {
                    // rare: the reference's FP64 expressions, from the stored chroma samples
                    const int yi[4] = {y0, y1, y2, y3};
                    const int16_t* pcb = reinterpret_cast<const int16_t*>(&s_cb[(ry >> 1) * kICStride + (mcu * 8 + c0) * 2]);
                    const int16_t* pcr = reinterpret_cast<const int16_t*>(&s_cr[(ry >> 1) * kICStride + (mcu * 8 + c0) * 2]);
                    ro[q4] = colour_exact4(yi, pcb[0], pcr[0], pcb[1], pcr[1], 0);
                    go[q4] = colour_exact4(yi, pcb[0], pcr[0], pcb[1], pcr[1], 1);
                    bo[q4] = colour_exact4(yi, pcb[0], pcr[0], pcb[1], pcr[1], 2);
                } else {
                    const int fr0 = sx_lo(oa[c0]), fg0 = sx_hi(oa[c0]), fb0 = sx_lo(ob[c0]);
                    const int fr1 = sx_lo(oa[c1]), fg1 = sx_hi(oa[c1]), fb1 = sx_lo(ob[c1]);
                    ro[q4] = pack_sat_u8x4(y0 + fr0, y1 + fr0, y2 + fr1, y3 + fr1);
                    go[q4] = pack_sat_u8x4(y0 + fg0, y1 + fg0, y2 + fg1, y3 + fg1);
                    bo[q4] = pack_sat_u8x4(y0 + fb0, y1 + fb0, y2 + fb1, y3 + fb1);
                }
            }
        }
        const size_t rowoff = (size_t(my) * 16 + ry) * p.W;
        if ((p.W & 15u) == 0 && x0 + 16u <= p.W) {
            *reinterpret_cast<uint4*>(R + rowoff + x0) = make_uint4(ro[0], ro[1], ro[2], ro[3]);
            *reinterpret_cast<uint4*>(G + rowoff + x0) = make_uint4(go[0], go[1], go[2], go[3]);
            *reinterpret_cast<uint4*>(B + rowoff + x0) = make_uint4(bo[0], bo[1], bo[2], bo[3]);
        } else {
#pragma unroll 1
            for (int i = 0; i < 16; ++i) {
                if (x0 + i >= p.W) break;
                R[rowoff + x0 + i] = uint8_t(ro[i >> 2] >> (8 * (i & 3)));
                G[rowoff + x0 + i] = uint8_t(go[i >> 2] >> (8 * (i & 3)));
                B[rowoff + x0 + i] = uint8_t(bo[i >> 2] >> (8 * (i & 3)));
            }
        }
    }
}

}  // namespace jz
