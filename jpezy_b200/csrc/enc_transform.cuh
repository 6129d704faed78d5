// enc_transform.cuh -- stages E1+E2: planar RGB -> quantised zig-zag int16 coefficients.
//
// Replaces make_YCC (src/encoder/jpezy_encoder.hpp:90-144), RGB::Y/Cb/Cr (:244-256), DCT (:146-166)
// and quantization (:168-172) of the reference.  One CTA transforms a strip of kMcuPerCta MCUs of
// one MCU row: colour conversion + 2:1 decimation into shared-memory MCU tiles, then the 8x8 DCTs,
// quantisation, zig-zag and a coalesced store in scan order (Y0 Y1 Y2 Y3 Cb Cr per MCU).
//
// Numerics: the reference truncates FP64 expressions, so results flip on 1-ulp differences when
// the exact value sits on a boundary.  The fast path computes the DCT with a separable transform;
// any coefficient whose value lies within kGuard of a non-zero multiple of its quantiser is
// recomputed in the reference's exact operation order (dct_exact), so the output is
// bit-identical to the reference's strict-IEEE evaluation.  Algorithmic HBM traffic: 3 B/px read
// + 3 B/px written (1.5 int16 coefficients per pixel) = 6 B/px.
#pragma once
#include "common.cuh"

namespace jz {

constexpr int kMcuPerCta = 8;                 // 128 x 16 pixels per CTA
constexpr int kBlkPerCta = kMcuPerCta * 6;    // 48 8x8 blocks
constexpr int kFwdThreads = 256;

struct FwdParams {
    const uint8_t *r, *g, *b;   // planes of image 0 (of the local rows when sharded)
    size_t plane_stride;        // bytes between consecutive images
    int16_t* coefs;
    size_t coef_stride;         // int16 elements between consecutive images
    uint32_t W, H;              // image size (H = full image height, used for edge replication)
    uint32_t HU, VU;            // MCUs per row / MCU rows handled by this launch
    uint32_t row0;              // first MCU row of this launch within the image (shards)
    uint32_t y_origin;          // image row stored at plane offset 0 (shards)
    int gray;
    unsigned long long* guard_counter;
};

// ---- colour conversion, bit-exact with src/encoder/jpezy_encoder.hpp:245-256 ------------------
// The reference evaluates left to right in double and truncates toward zero.
__device__ __forceinline__ int ref_Y(int r, int g, int b)
{
    double t = __dadd_rn(__dmul_rn(0.2990, double(r)), __dmul_rn(0.5870, double(g)));
    t = __dadd_rn(t, __dmul_rn(0.1140, double(b)));
    return __double2int_rz(__dadd_rn(t, -128.0));
}
__device__ __forceinline__ int ref_Cb(int r, int g, int b)
{
    double t = __dsub_rn(-__dmul_rn(0.1687, double(r)), __dmul_rn(0.3313, double(g)));
    return __double2int_rz(__dadd_rn(t, __dmul_rn(0.5000, double(b))));
}
__device__ __forceinline__ int ref_Cr(int r, int g, int b)
{
    double t = __dsub_rn(__dmul_rn(0.5000, double(r)), __dmul_rn(0.4187, double(g)));
    return __double2int_rz(__dsub_rn(t, __dmul_rn(0.0813, double(b))));
}

// Integer evaluation of the same formulas: the exact value is n/10000 with n integer, and the
// FP64 evaluation is within 1e-12 of it, so truncation agrees unless n is a multiple of 10000
// (exact integer result), where the FP64 rounding decides and ref_* is used instead.
__device__ __forceinline__ int fast_Y(int r, int g, int b)
{
    const int n = 2990 * r + 5870 * g + 1140 * b - 1280000;   // |n| <= 1.28e6
    const int q = n / 10000;
    return (n - q * 10000 == 0) ? ref_Y(r, g, b) : q;
}
__device__ __forceinline__ int fast_Cb(int r, int g, int b)
{
    const int n = -1687 * r - 3313 * g + 5000 * b;
    const int q = n / 10000;
    return (n - q * 10000 == 0) ? ref_Cb(r, g, b) : q;
}
__device__ __forceinline__ int fast_Cr(int r, int g, int b)
{
    const int n = 5000 * r - 4187 * g - 813 * b;
    const int q = n / 10000;
    return (n - q * 10000 == 0) ? ref_Cr(r, g, b) : q;
}

// ---- exact-order DCT of one coefficient (src/encoder/jpezy_encoder.hpp:146-166) ----------------
// F[i][j] = int( (sum_y sum_x ((pic*cos[j][x])*cos[i][y])) * cu * cv / 4 ), y outer, x inner.
__device__ __noinline__ double dct_exact(const int8_t* __restrict__ pic, int i, int j)
{
    double sum = 0.0;
    for (int y = 0; y < 8; ++y) {
        const double ci = cC.cos_ref[i * 8 + y];
#pragma unroll
        for (int x = 0; x < 8; ++x)
            sum = __dadd_rn(sum, __dmul_rn(__dmul_rn(double(pic[y * 8 + x]), cC.cos_ref[j * 8 + x]), ci));
    }
    const double cu = j ? 1.0 : cC.inv_sqrt2_ref;
    const double cv = i ? 1.0 : cC.inv_sqrt2_ref;
    return __dmul_rn(__dmul_rn(__dmul_rn(sum, cu), cv), 0.25);   // "/ 4" is exact
}

constexpr double kGuardF64 = 1e-6;   // fast-path FP64 error is < 1e-10; see DESIGN.md

// ---- validation build: FP64 separable DCT ---------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads) k_fwd_transform_f64(const FwdParams p)
{
    __shared__ __align__(16) int8_t s_pix[kBlkPerCta][64];
    __shared__ double s_tmp[kBlkPerCta][64];
    __shared__ __align__(16) int16_t s_out[kBlkPerCta][64];

    const int t = threadIdx.x;
    const uint32_t mx0 = blockIdx.x * kMcuPerCta;
    const uint32_t my = blockIdx.y;
    const size_t img = blockIdx.z;
    const uint8_t* __restrict__ R = p.r + img * p.plane_stride;
    const uint8_t* __restrict__ G = p.g + img * p.plane_stride;
    const uint8_t* __restrict__ B = p.b + img * p.plane_stride;

    // ---- phase 1: colour conversion + decimation into MCU tiles ----
    {
        const int pr = t >> 5;    // row pair 0..7
        const int c4 = t & 31;    // group of 4 pixels
        const int mcu = c4 >> 2;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int ry = pr * 2 + rr;
            uint32_t gy = (p.row0 + my) * 16u + ry;
            if (gy > p.H - 1) gy = p.H - 1;
            const size_t rowoff = size_t(gy - p.y_origin) * p.W;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cx = c4 * 4 + i;
                uint32_t gx = mx0 * 16u + cx;
                if (gx > p.W - 1) gx = p.W - 1;
                const int rv = __ldg(R + rowoff + gx), gv = __ldg(G + rowoff + gx), bv = __ldg(B + rowoff + gx);
                const int k = (ry >> 3) * 2 + ((cx & 15) >> 3);
                s_pix[mcu * 6 + k][(ry & 7) * 8 + (cx & 7)] = int8_t(fast_Y(rv, gv, bv));
                if (rr == 0 && (i & 1) == 0) {
                    const int pos = pr * 8 + ((cx & 15) >> 1);
                    s_pix[mcu * 6 + 4][pos] = p.gray ? int8_t(0) : int8_t(fast_Cb(rv, gv, bv));
                    s_pix[mcu * 6 + 5][pos] = p.gray ? int8_t(0) : int8_t(fast_Cr(rv, gv, bv));
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2a: row pass  tmp[y][j] = sum_x pic[y][x] cos[j][x] ----
    for (int task = t; task < kBlkPerCta * 8; task += kFwdThreads) {
        const int blk = task >> 3, y = task & 7;
        double px[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) px[x] = double(s_pix[blk][y * 8 + x]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double s = 0.0;
#pragma unroll
            for (int x = 0; x < 8; ++x) s = fma(px[x], cC.cos_ref[j * 8 + x], s);
            s_tmp[blk][y * 8 + j] = s;
        }
    }
    __syncthreads();

    // ---- phase 2b: column pass + quantisation + zig-zag ----
    unsigned long long guard_hits = 0;
    for (int task = t; task < kBlkPerCta * 8; task += kFwdThreads) {
        const int blk = task >> 3, j = task & 7;
        const int cs = (blk % 6) >= 4;
        double col[8];
#pragma unroll
        for (int y = 0; y < 8; ++y) col[y] = s_tmp[blk][y * 8 + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double s = 0.0;
#pragma unroll
            for (int y = 0; y < 8; ++y) s = fma(col[y], cC.cos_ref[i * 8 + y], s);
            const double scale = 0.25 * (i ? 1.0 : 0.70710678118654752440) * (j ? 1.0 : 0.70710678118654752440);
            double v = s * scale;
            const int q = cC.quant[cs][i * 8 + j];
            const double k = rint(v / double(q));
            if (k != 0.0 && fabs(v - k * double(q)) < kGuardF64) {
                v = dct_exact(&s_pix[blk][0], i, j);
                ++guard_hits;
            }
            s_out[blk][cC.izz[i * 8 + j]] = int16_t(__double2int_rz(v) / q);
        }
    }
    __syncthreads();
    if (guard_hits) atomicAdd(p.guard_counter, guard_hits);

    // ---- phase 3: coalesced store, scan order ----
    const uint32_t nvalid = min(uint32_t(kMcuPerCta), p.HU - mx0);
    const uint32_t nwords = nvalid * 6 * 32;   // uint32 words
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.coefs + img * p.coef_stride + (size_t(my) * p.HU + mx0) * 384);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&s_out[0][0]);
    for (uint32_t w = t; w < nwords; w += kFwdThreads) dst[w] = src[w];
}

}  // namespace jz
