// dec_transform.cuh -- stages D2+D3: quantised coefficients -> planar RGB.
//
// Replaces inverse_quantization (src/decoder/jpezy_decoder.hpp:645-650), inverse_dct (:652-670), the
// pixel replication of decode_mcu (:519-524) and make_rgb / to_r,g,b / revise_value (:531-578, :672-676).
// One CTA reconstructs a strip of kMcuPerCta MCUs: dequantise, 8x8 IDCT (separable fast path), chroma
// replicated 2x2 from shared memory, YCbCr->RGB, planar stores with row stride = image width.
//
// Numerics: the reference truncates int(sum/4 + 128) of an FP64 sum, and the exact value is very often
// an integer (every DC-only block with an even DC*q/8), where the reference's rounding decides.  Samples
// whose fast-path value is within kGuardInv of an integer are recomputed in the reference's exact
// operation order over the block's non-zero coefficients (zero terms add +-0 and cannot change the sum).
// Algorithmic HBM traffic: 3 B/px read (int16 coefficients) + 3 B/px written = 6 B/px.
#pragma once
#include "common.cuh"
#include "enc_transform.cuh"

namespace jz {

struct InvParams {
    const int16_t* coefs;
    size_t coef_stride;
    uint8_t *r, *g, *b;
    size_t plane_stride;      // bytes between images (= plane_bytes)
    uint32_t W, H, HU, VU;
    int gray;
    uint16_t qt[3][64];       // per component, natural order
    unsigned long long* guard_counter;
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

__global__ void __launch_bounds__(kFwdThreads) k_inv_transform_f64(const InvParams p)
{
    __shared__ int s_dq[kBlkPerCta][64];                 // dequantised, natural order
    __shared__ double s_tmp[kBlkPerCta][64];
    __shared__ short s_pix[kBlkPerCta][64];              // IDCT output (unclamped int, fits 16 bits)
    __shared__ unsigned long long s_nz[kBlkPerCta];

    const int t = threadIdx.x;
    const uint32_t mx0 = blockIdx.x * kMcuPerCta;
    const uint32_t my = blockIdx.y;
    const size_t img = blockIdx.z;
    const uint32_t nvalid = min(uint32_t(kMcuPerCta), p.HU - mx0);

    // ---- load + dequantise (zig-zag -> natural) ----
    const int16_t* src = p.coefs + img * p.coef_stride + (size_t(my) * p.HU + mx0) * 384;
    for (uint32_t e = t; e < kBlkPerCta * 64; e += kFwdThreads) {
        const uint32_t blk = e >> 6, n = e & 63;
        const int comp = (blk % 6) < 4 ? 0 : int(blk % 6) - 3;
        const int nat = cC.zz[n];
        const int c = (blk / 6) < nvalid ? int(src[e]) : 0;
        s_dq[blk][nat] = c * int(p.qt[comp][nat]);
    }
    __syncthreads();
    if (t < kBlkPerCta) {
        unsigned long long m = 0;
        for (int i = 0; i < 64; ++i)
            if (s_dq[t][i]) m |= 1ull << i;
        s_nz[t] = m;
    }

    // ---- IDCT pass 1 (over v): tmp[y][u] = sum_v cv * F[v][u] * cos[v][y] ----
    for (int task = t; task < kBlkPerCta * 8; task += kFwdThreads) {
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

    // ---- IDCT pass 2 (over u) + level shift + guard ----
    unsigned long long guard_hits = 0;
    for (int task = t; task < kBlkPerCta * 8; task += kFwdThreads) {
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

    // ---- upsample + colour + store ----
    uint8_t* R = p.r + img * p.plane_stride;
    uint8_t* G = p.g + img * p.plane_stride;
    uint8_t* B = p.b + img * p.plane_stride;
    for (int e = t; e < 16 * 128; e += kFwdThreads) {
        const int ry = e >> 7, cx = e & 127;
        const uint32_t gx = mx0 * 16u + cx;
        if (gx >= p.W) continue;
        const int mcu = cx >> 4;
        const int k = (ry >> 3) * 2 + ((cx & 15) >> 3);
        const int yv = s_pix[mcu * 6 + k][(ry & 7) * 8 + (cx & 7)];
        const size_t idx = (size_t(my) * 16 + ry) * p.W + gx;
        if (p.gray) {
            const uint8_t v = revise(double(yv));
            R[idx] = v, G[idx] = v, B[idx] = v;
        } else {
            const int cpos = (ry >> 1) * 8 + ((cx & 15) >> 1);
            const int cb = s_pix[mcu * 6 + 4][cpos], cr = s_pix[mcu * 6 + 5][cpos];
            R[idx] = ref_R(yv, cr), G[idx] = ref_G(yv, cb, cr), B[idx] = ref_B(yv, cb);
        }
    }
}

}  // namespace jz
