// capi.cu -- the C ABI of libjpezy_b200.so (include/jpezy_b200.h): context, table upload,
// kernel launches.  Single translation unit: all kernels are included here.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "dec_entropy.cuh"
#include "dec_transform.cuh"
#include "dec_transform2.cuh"
#include "dec_transform_g.cuh"
#include "enc_entropy.cuh"
#include "enc_transform.cuh"
#include "enc_transform2.cuh"
#include "enc_shard.cuh"
#include "synth.cuh"

namespace jz {

// canonical Huffman code construction (T.81 Annex C.2); the same thing the reference's decoder
// does in analyze_dht (src/decoder/jpezy_decoder.hpp:223-239)
static void canonical_codes(const uint8_t bits[16], const uint8_t* vals, int nvals, uint16_t code_of[256], uint8_t len_of[256])
{
    std::memset(code_of, 0, 256 * sizeof(uint16_t));
    std::memset(len_of, 0, 256);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int j = 0; j < bits[len - 1] && k < nvals; ++j, ++k) {
            code_of[vals[k]] = uint16_t(code++);
            len_of[vals[k]] = uint8_t(len);
        }
        code <<= 1;
    }
}

static void build_enc_lut(const HuffSpec& dc, const HuffSpec& ac, HuffEncLut* out)
{
    uint16_t code[256];
    uint8_t len[256];
    std::memset(out, 0, sizeof *out);
    canonical_codes(dc.bits, dc.vals, dc.nvals, code, len);
    for (int i = 0; i < 16; ++i) out->dc[i] = len[i] ? (uint32_t(code[i]) << 5) | len[i] : 0u;
    canonical_codes(ac.bits, ac.vals, ac.nvals, code, len);
    for (int i = 0; i < 256; ++i) out->ac[i] = len[i] ? (uint32_t(code[i]) << 5) | len[i] : 0u;
}

static cudaStream_t pick_stream(jpezyb200_ctx* ctx, void* s) { return s ? static_cast<cudaStream_t>(s) : ctx->stream; }
static void batch_pipe_destroy(jpezyb200_ctx* ctx);

}  // namespace jz

using namespace jz;

extern "C" {

static void host_pipe_destroy(jpezyb200_ctx* ctx);

int jpezyb200_abi_version(void) { return JPEZYB200_ABI_VERSION; }

const char* jpezyb200_strerror(int code)
{
    switch (code) {
    case JPEZYB200_OK: return "ok";
    case JPEZYB200_EINVAL: return "invalid argument";
    case JPEZYB200_ECAPACITY: return "output buffer too small";
    case JPEZYB200_ECUDA: return "CUDA error";
    case JPEZYB200_ENCCL: return "collective error";
    case JPEZYB200_ECORRUPT: return "corrupt entropy-coded segment";
    case JPEZYB200_ENODEVICE: return "no CUDA device (libjpezy_b200 has no CPU path)";
    case JPEZYB200_ENOMEM: return "out of memory";
    case JPEZYB200_EUNSUPPORTED: return "unsupported frame layout";
    case JPEZYB200_EAGAIN: return "parallel Huffman decoder needs more synchronisation rounds";
    default: return "unknown error";
    }
}

const char* jpezyb200_last_error(const jpezyb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int jpezyb200_ctx_create(int device, jpezyb200_ctx** out)
{
    if (!out) return JPEZYB200_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return JPEZYB200_ENODEVICE;
    }
    if (device < 0 || device >= ndev) return JPEZYB200_EINVAL;
    jpezyb200_ctx* ctx = new (std::nothrow) jpezyb200_ctx();
    if (!ctx) return JPEZYB200_ENOMEM;
    ctx->device = device;
    if (cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->num_sms < 1) ctx->num_sms = 148;
    if (const char* e = std::getenv("JPEZY_B200_SYNC_ROUNDS")) ctx->sync_rounds = std::atoi(e);
    if (const char* e = std::getenv("JPEZY_B200_DEC_HYP")) ctx->sync_guesses = std::max(0, std::min(2, std::atoi(e)));      // (JPEZYB200_OPT_SYNC_GUESSES of every new context)      // (debugging: JPEZYB200_OPT_SYNC_ROUNDS of every new context)
    int rc = [&]() -> int {
        JZ_CUDA_TRY(ctx, cudaSetDevice(device));
        JZ_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        DevConst h{};
        for (int i = 0; i < 64; ++i) {
            h.cos_ref[i] = kCosRef[i];
            h.cosf_[i] = float(std::cos((2 * (i & 7) + 1) * (i >> 3) * 3.14159265358979323846 / 16));
            h.quant[0][i] = kQuantLuma[i], h.quant[1][i] = kQuantChroma[i];
            h.rquant[0][i] = 1.0f / float(kQuantLuma[i]), h.rquant[1][i] = 1.0f / float(kQuantChroma[i]);
            h.zz[i] = kZigzag[i];
            h.izz[kZigzag[i]] = uint8_t(i);
        }
        h.inv_sqrt2_ref = kInvSqrt2Ref;
        JZ_CUDA_TRY(ctx, cudaMemcpyToSymbol(cC, &h, sizeof h));
        JZ_CUDA_TRY(ctx, cudaMemcpyToSymbol(gCosRef, h.cos_ref, sizeof h.cos_ref));
        {
            // constants of the FP32 AAN path: K folds the AAN scale factors, the /8 and 1/q; G is the guard band
            // (1.25 x worst-case flowgraph error + the rounding of w itself) in w units; T the "quantises to 0" test
            QuantConst qc{};
            double aan[8];
            aan[0] = 1.0;
            for (int k = 1; k < 8; ++k) aan[k] = std::cos(k * 3.14159265358979323846 / 16) * std::sqrt(2.0);
            for (int c = 0; c < 2; ++c)
                for (int i = 0; i < 8; ++i)
                    for (int j = 0; j < 8; ++j) {
                        const double q = c ? kQuantChroma[i * 8 + j] : kQuantLuma[i * 8 + j];
                        const double K = 1.0 / (8.0 * aan[i] * aan[j] * q);
                        const double Gd = (1.25 * kAanErrBound[i * 8 + j] + 5e-4) / q;
                        qc.K[c][i * 8 + j] = float(K);
                        qc.G[c][i * 8 + j] = float(Gd);
                        qc.T[c][i * 8 + j] = float((1.0 - 2.0 * Gd) / K);
                    }
            JZ_CUDA_TRY(ctx, cudaMemcpyToSymbol(cQ, &qc, sizeof qc));
            QuantCol col[2][8];
            uint2 izzc[8];
            for (int j = 0; j < 8; ++j) {
                uint8_t zb[8];
                for (int i = 0; i < 8; ++i) {
                    for (int c = 0; c < 2; ++c) col[c][j].K[i] = qc.K[c][i * 8 + j], col[c][j].T[i] = qc.T[c][i * 8 + j], col[c][j].G[i] = qc.G[c][i * 8 + j];
                    zb[i] = h.izz[i * 8 + j];
                }
                izzc[j].x = uint32_t(zb[0]) | (uint32_t(zb[1]) << 8) | (uint32_t(zb[2]) << 16) | (uint32_t(zb[3]) << 24);
                izzc[j].y = uint32_t(zb[4]) | (uint32_t(zb[5]) << 8) | (uint32_t(zb[6]) << 16) | (uint32_t(zb[7]) << 24);
            }
            JZ_CUDA_TRY(ctx, cudaMemcpyToSymbol(gQcol, col, sizeof col));
            JZ_CUDA_TRY(ctx, cudaMemcpyToSymbol(gIzzCol, izzc, sizeof izzc));
            // second-generation forward kernel (enc_transform2.cuh): (K, K) pairs per (class, column); the DC multiplier
            // carries 1 - 2^-20 so that trunc(S * K) reproduces int(int(((S*c)*c)/4)/q) for every sum S; class 2 = zeros
            Q2Tab q2;
            std::memset(&q2, 0, sizeof q2);
            for (int c = 0; c < 2; ++c) {
                double gmax = 0.0;
                for (int j = 0; j < 8; ++j)
                    for (int i = 0; i < 8; ++i) {
                        float K = qc.K[c][i * 8 + j];
                        if ((i | j) == 0) K = float((1.0 - 1.0 / 1048576.0) / (8.0 * (c ? kQuantChroma[0] : kQuantLuma[0])));
                        else gmax = std::max(gmax, double(qc.G[c][i * 8 + j]));
                        q2.K[c][i >> 1][j][i & 1] = make_float2(K, K);
                        q2.G[c][i][j] = qc.G[c][i * 8 + j];
                    }
                const float thr = float(1.0 - 2.0 * gmax);
                std::memcpy(&q2.thr[c], &thr, 4);
            }
            for (int j = 0; j < 8; ++j) q2.izz[j] = izzc[j];
            JZ_CUDA_TRY(ctx, cudaMemcpyToSymbol(gQ2, &q2, sizeof q2));
        }
        HuffEncLut lut[2];
        build_enc_lut(kDcLuma, kAcLuma, &lut[0]);
        build_enc_lut(kDcChroma, kAcChroma, &lut[1]);
        JZ_CUDA_TRY(ctx, cudaMalloc(&ctx->d_enc_lut, sizeof lut));
        JZ_CUDA_TRY(ctx, cudaMemcpy(ctx->d_enc_lut, lut, sizeof lut, cudaMemcpyHostToDevice));
        JZ_CUDA_TRY(ctx, cudaMalloc(&ctx->d_y_exact, 65536));
        k_build_y_exact<<<256, 256, 0, ctx->stream>>>(ctx->d_y_exact);
        JZ_CUDA_TRY(ctx, cudaGetLastError());
        JZ_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        JZ_CUDA_TRY(ctx, cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
        JZ_CUDA_TRY(ctx, cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long)));
        return JPEZYB200_OK;
    }();
    if (rc != JPEZYB200_OK) {
        jpezyb200_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return JPEZYB200_OK;
}

void jpezyb200_ctx_destroy(jpezyb200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    jz_devbuf* bufs[] = {&ctx->coefs, &ctx->blk_off, &ctx->tile_sum, &ctx->tile_base, &ctx->img_bits, &ctx->ustream,
                         &ctx->ff_sum, &ctx->ff_base, &ctx->planes_in, &ctx->planes_out, &ctx->scan_io, &ctx->sizes_io,
                         &ctx->dec_scanbytes, &ctx->dec_chunk_cnt, &ctx->dec_chunk_base, &ctx->dec_ubytes, &ctx->dec_state,
                         &ctx->dec_dirty, &ctx->dec_subblk, &ctx->dec_dc, &ctx->dec_dcd, &ctx->blk_meta, &ctx->dec_status, &ctx->dec_changed, &ctx->dec_flags, &ctx->inv_samples, &ctx->shard_geom, &ctx->dec_mcnt, &ctx->dec_mbase, &ctx->dec_seg};
    for (jz_devbuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->d_enc_lut) cudaFree(ctx->d_enc_lut);
    if (ctx->d_dec_lut) cudaFree(ctx->d_dec_lut);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_y_exact) cudaFree(ctx->d_y_exact);
    for (void* tb : ctx->invg_tab)
        if (tb) cudaFree(tb);
    for (void* tb : ctx->inv2_tab)
        if (tb) cudaFree(tb);
    jz::batch_pipe_destroy(ctx);
    host_pipe_destroy(ctx);
    std::free(ctx->shard_state);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_sizes) cudaFreeHost(ctx->h_sizes);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream), cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
    if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int jpezyb200_set_option(jpezyb200_ctx* ctx, int option, int64_t value)
{
    if (!ctx) return JPEZYB200_EINVAL;
    switch (option) {
    case JPEZYB200_OPT_PAD_ONES: ctx->pad_ones = value ? 1 : 0; return JPEZYB200_OK;
    case JPEZYB200_OPT_TRANSFORM: ctx->transform_variant = int(value); return JPEZYB200_OK;
    case JPEZYB200_OPT_BATCH_GROUP_BYTES:
        if (value < 1) return ctx->fail(JPEZYB200_EINVAL, "group bytes must be positive");
        ctx->group_bytes = value;
        return JPEZYB200_OK;
    case JPEZYB200_OPT_SHARD_SCRATCH_BYTES:
        if (value < 0) return ctx->fail(JPEZYB200_EINVAL, "scratch bytes must be >= 0");
        ctx->shard_scratch = value;
        return JPEZYB200_OK;
    case JPEZYB200_OPT_SYNC_GUESSES:
        if (value < 0 || value > 2) return ctx->fail(JPEZYB200_EINVAL, "sync guesses: 0 (one per subsequence), 1 (one per block position on latency-bound inputs), 2 (always)");
        ctx->sync_guesses = int(value);
        return JPEZYB200_OK;
    case JPEZYB200_OPT_SYNC_ROUNDS:
        if (value < 0 || value > 64) return ctx->fail(JPEZYB200_EINVAL, "sync rounds must be in 0..64");
        ctx->sync_rounds = int(value);
        return JPEZYB200_OK;
    default: return ctx->fail(JPEZYB200_EINVAL, "unknown option");
    }
}

int jpezyb200_get_stat(jpezyb200_ctx* ctx, int stat, uint64_t* value)
{
    if (!ctx || !value) return JPEZYB200_EINVAL;
    if (stat == JPEZYB200_STAT_KERNEL_LAUNCHES) {
        *value = ctx->launches;
        return JPEZYB200_OK;
    }
    int idx = stat == JPEZYB200_STAT_GUARD_FWD ? 0 : stat == JPEZYB200_STAT_GUARD_INV ? 1 : stat == JPEZYB200_STAT_SYNC_ROUNDS ? 2 :
              stat == JPEZYB200_STAT_SYNC_ITERS0 ? 4 : stat == JPEZYB200_STAT_SYNC_ITERS1 ? 5 : -1;
    if (idx < 0) return ctx->fail(JPEZYB200_EINVAL, "unknown stat");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    unsigned long long v = 0;
    JZ_CUDA_TRY(ctx, cudaMemcpy(&v, ctx->d_counters + idx, sizeof v, cudaMemcpyDeviceToHost));
    *value = v;
    return JPEZYB200_OK;
}

// -------------------------------------------------------------------------------------------------
// encoder
// -------------------------------------------------------------------------------------------------
// copy stream + events of the band-pipelined single-image host entry points (jpezyb200_encode / jpezyb200_decode)
constexpr int kHostBands = 8;
// bands of MCU rows whose copies overlap the transform kernels.  Measured on a 4K frame: 1 band 6.48 GPix/s, 2: 6.49, 3: 6.44,
// 4: 6.34, 8: 6.23 -- the round trip is bound by the two 25 MB PCIe copies, the 77 us of transform time that could hide behind
// them do not pay for the extra copy calls.  Default 1 (JPEZY_B200_HOST_BANDS overrides, for experiments).
static size_t host_bands()
{
    static const size_t n = [] { const char* e = std::getenv("JPEZY_B200_HOST_BANDS"); const long v = e ? std::atol(e) : 1; return size_t(v < 1 ? 1 : (v > kHostBands ? kHostBands : v)); }();
    return n;
}
struct HostPipe {
    cudaStream_t copy = nullptr;
    cudaEvent_t ev[kHostBands] = {};
    uint64_t* h_sz = nullptr;      // pinned: sizes / status read back by the host
};
static void host_pipe_destroy(jpezyb200_ctx* ctx);
static int host_pipe(jpezyb200_ctx* ctx, HostPipe** out)
{
    if (!ctx->host_pipe) {
        // built completely before it is used: a half-made pipe (a failed stream / event / pinned allocation) is torn down again
        HostPipe* hp = new (std::nothrow) HostPipe();
        if (!hp) return JPEZYB200_ENOMEM;
        ctx->host_pipe = hp;
        const int rc = [&]() -> int {
            JZ_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&hp->copy, cudaStreamNonBlocking));
            for (cudaEvent_t& e : hp->ev) JZ_CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            JZ_CUDA_TRY(ctx, cudaMallocHost(&hp->h_sz, 64));
            return JPEZYB200_OK;
        }();
        if (rc != JPEZYB200_OK) {
            host_pipe_destroy(ctx);
            return rc;
        }
    }
    *out = static_cast<HostPipe*>(ctx->host_pipe);
    return JPEZYB200_OK;
}
static void host_pipe_destroy(jpezyb200_ctx* ctx)
{
    HostPipe* hp = static_cast<HostPipe*>(ctx->host_pipe);
    if (!hp) return;
    if (hp->copy) cudaStreamSynchronize(hp->copy), cudaStreamDestroy(hp->copy);
    for (cudaEvent_t e : hp->ev)
        if (e) cudaEventDestroy(e);
    if (hp->h_sz) cudaFreeHost(hp->h_sz);
    delete hp;
    ctx->host_pipe = nullptr;
}

static int check_geometry(jpezyb200_ctx* ctx, uint32_t W, uint32_t H, uint32_t nimg)
{
    if (!ctx) return JPEZYB200_EINVAL;
    // SOF0 carries 16-bit sizes (src/encoder/jpezy_writer.hpp:70-71)
    if (W == 0 || H == 0 || W > 65535u || H > 65535u) return ctx->fail(JPEZYB200_EINVAL, "width/height must be in 1..65535");
    if (nimg == 0 || nimg > 65535u) return ctx->fail(JPEZYB200_EINVAL, "nimg must be in 1..65535");
    return JPEZYB200_OK;
}

// the bulk copies of k_fwd_transform2 move whole 16-byte units: every pixel row and the coefficient buffer must start on one
static bool fwd2_ok(const FwdParams& p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.r) | reinterpret_cast<uintptr_t>(p.g) | reinterpret_cast<uintptr_t>(p.b) |
                        reinterpret_cast<uintptr_t>(p.coefs);
    return (p.W & 15u) == 0 && (a & 15u) == 0 && (p.plane_stride & 15u) == 0;
}

// picks the forward-transform kernel for the parameters (JPEZYB200_OPT_TRANSFORM, alignment) and launches it
static int launch_fwd_kernel(jpezyb200_ctx* ctx, const FwdParams& p, uint32_t nimg, cudaStream_t st)
{
    if (ctx->transform_variant == 1) {
        dim3 grid((p.HU + kMcuPerCta - 1) / kMcuPerCta, p.VU, nimg);
        k_fwd_transform_f64<<<grid, kFwdThreads, 0, st>>>(p);
    } else if (ctx->transform_variant == 3 && fwd2_ok(p)) {
        // second-generation kernel: persistent CTAs, rows fetched by bulk copies (needs 16-byte aligned rows and buffers).  Measured
        // against the production kernel below in one run: equal on a single 4K frame (28.3 us), 10 % slower on batches
        // (404 against 365 us for 64 HD frames), half as fast on S-noise -- so it stays an A/B variant (JPEZYB200_OPT_TRANSFORM = 3)
        // JPEZY_B200_FWD_CFG (tuning runs): 83 = the compute warps store their own blocks (default), 82 = the DMA warp stores whole tiles,
        // 84 = as 83 with four input stages
        static const int cfg = [] { const char* e = std::getenv("JPEZY_B200_FWD_CFG"); return e ? std::atoi(e) : 83; }();
        if (!ctx->fwd2_attr_set) {
            JZ_CUDA_TRY(ctx, cudaFuncSetAttribute(k_fwd_transform2<8, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2<8, 3>::kSmem));
            JZ_CUDA_TRY(ctx, cudaFuncSetAttribute(k_fwd_transform2<8, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2<8, 3>::kSmem));
            JZ_CUDA_TRY(ctx, cudaFuncSetAttribute(k_fwd_transform2<8, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2<8, 4>::kSmem));
            JZ_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->fwd2_occ[0], k_fwd_transform2<8, 3, true>, Fwd2<8, 3>::kThreads, Fwd2<8, 3>::kSmem));
            JZ_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->fwd2_occ[1], k_fwd_transform2<8, 3, false>, Fwd2<8, 3>::kThreads, Fwd2<8, 3>::kSmem));
            JZ_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->fwd2_occ[2], k_fwd_transform2<8, 4, true>, Fwd2<8, 4>::kThreads, Fwd2<8, 4>::kSmem));
            JZ_CUDA_TRY(ctx, cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, ctx->device));
            ctx->fwd2_attr_set = true;
        }
        const int T = 8, v = cfg == 82 ? 1 : (cfg == 84 ? 2 : 0);
        const uint32_t tpr = (p.HU + T - 1) / T;
        const uint64_t ntiles64 = uint64_t(tpr) * p.VU * nimg;
        if (ntiles64 > 0xffffffffull) return ctx->fail(JPEZYB200_EINVAL, "too many tiles");
        const uint32_t ntiles = uint32_t(ntiles64);
        static const int per_sm_cap = [] { const char* e = std::getenv("JPEZY_B200_FWD_CTAS"); return e ? std::atoi(e) : 0; }();
        const int per_sm = per_sm_cap > 0 ? std::min(per_sm_cap, ctx->fwd2_occ[v]) : ctx->fwd2_occ[v];
        const uint32_t grid = std::min<uint32_t>(ntiles, uint32_t(ctx->num_sms * std::max(1, per_sm)));
        // how (image, MCU row, tile of the row) advance when the tile id grows by the grid size
        TileStep ts{};
        ts.tiles_per_row = tpr;
        const uint32_t per_img = tpr * p.VU;
        ts.dimg = grid / per_img;
        ts.dmy = (grid % per_img) / tpr;
        ts.dbx = (grid % per_img) % tpr;
        static const uint32_t tsflags = [] { const char* e = std::getenv("JPEZY_B200_FWD_FLAGS"); return e ? uint32_t(std::atoi(e)) : 0u; }();
        ts.flags = tsflags;
        if (v == 0) (void)jz_launch(k_fwd_transform2<8, 3, true>, dim3(grid), dim3(Fwd2<8, 3>::kThreads), Fwd2<8, 3>::kSmem, st, p, ntiles, ts);
        else if (v == 1) (void)jz_launch(k_fwd_transform2<8, 3, false>, dim3(grid), dim3(Fwd2<8, 3>::kThreads), Fwd2<8, 3>::kSmem, st, p, ntiles, ts);
        else (void)jz_launch(k_fwd_transform2<8, 4, true>, dim3(grid), dim3(Fwd2<8, 4>::kThreads), Fwd2<8, 4>::kSmem, st, p, ntiles, ts);
    } else {
        dim3 grid((p.HU + kTileMcu - 1) / kTileMcu, p.VU, nimg);
        if (ctx->transform_variant == 2) (void)jz_launch(k_fwd_transform_t<false>, grid, dim3(256), 0, st, p);    // one thread per block (A/B runs)
        else {
            if (!ctx->fwd_attr_set) {   // static + dynamic shared memory exceed 48 KiB: opt in once per context
                JZ_CUDA_TRY(ctx, cudaFuncSetAttribute(k_fwd_transform_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdTrSmem));
                ctx->fwd_attr_set = true;
            }
            (void)jz_launch(k_fwd_transform_t<true>, grid, dim3(256), kFwdTrSmem, st, p);
        }
    }
    ++ctx->launches;
    JZ_CUDA_TRY(ctx, cudaGetLastError());
    return JPEZYB200_OK;
}

// rows: MCU rows [row0, row0 + nrows) only (nrows == 0: all); the planes always hold the whole image
static int launch_fwd(jpezyb200_ctx* ctx, const uint8_t* d_r, const uint8_t* d_g, const uint8_t* d_b, uint32_t W, uint32_t H,
                      uint32_t nimg, int gray, int16_t* d_coefs, cudaStream_t st, uint32_t row0 = 0, uint32_t nrows = 0,
                      bool with_meta = false)
{
    FwdParams p{};
    p.r = d_r, p.g = d_g, p.b = d_b;
    p.plane_stride = size_t(W) * H;
    p.W = W, p.H = H;
    p.HU = mcu_units(W), p.VU = nrows ? nrows : mcu_units(H);
    p.coefs = d_coefs + size_t(row0) * p.HU * 384;
    p.coef_stride = size_t(p.HU) * mcu_units(H) * 384;
    p.row0 = row0, p.y_origin = 0;
    p.gray = gray;
    p.guard_counter = ctx->d_counters + 0;
    p.y_exact = ctx->d_y_exact;
    if (with_meta && ctx->transform_variant != 1) {     // side information for launch_entropy (same context, same images)
        int rc;
        if ((rc = ctx->ensure(ctx->blk_meta, size_t(nimg) * (p.coef_stride >> 6) * 4))) return rc;
        p.bmeta = static_cast<uint32_t*>(ctx->blk_meta.p) + size_t(row0) * p.HU * 6;
    }
    return launch_fwd_kernel(ctx, p, nimg, st);
}

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int launch_entropy(jpezyb200_ctx* ctx, const int16_t* d_coefs, uint32_t W, uint32_t H, uint32_t nimg, uint8_t* d_scan,
                          size_t slot_bytes, uint64_t* d_scan_bytes, uint64_t* d_scan_bits, cudaStream_t st, bool with_meta = false)
{
    const uint32_t HU = mcu_units(W), VU = mcu_units(H);
    const size_t nmcu = size_t(HU) * VU;
    if (nmcu * 6 > 0xffffff00ull) return ctx->fail(JPEZYB200_EINVAL, "too many blocks");
    EntParams p{};
    p.coefs = d_coefs;
    p.coef_stride = nmcu * 384;
    p.bmeta = with_meta && ctx->transform_variant != 1 ? static_cast<const uint32_t*>(ctx->blk_meta.p) : nullptr;
    p.nblk = uint32_t(nmcu * 6);
    p.ntile = (p.nblk + kEntThreads - 1) / kEntThreads;
    p.uslot = round_up(slot_bytes + 64, 16);
    p.nchunk = uint32_t((p.uslot + kStuffChunk - 1) / kStuffChunk) + 1;
    p.slot = slot_bytes;
    p.out = d_scan;
    p.out_bytes = d_scan_bytes, p.out_bits = d_scan_bits;
    p.lut = ctx->d_enc_lut;
    p.dc_init = nullptr;
    p.pad_ones = ctx->pad_ones;
    int rc;
    if ((rc = ctx->ensure(ctx->blk_off, size_t(nimg) * p.nblk * 4))) return rc;
    if ((rc = ctx->ensure(ctx->tile_sum, size_t(nimg) * p.ntile * 4))) return rc;
    if ((rc = ctx->ensure(ctx->tile_base, size_t(nimg) * p.ntile * 8))) return rc;
    if ((rc = ctx->ensure(ctx->img_bits, size_t(nimg) * 16))) return rc;
    if ((rc = ctx->ensure(ctx->ustream, size_t(nimg) * p.uslot))) return rc;
    if ((rc = ctx->ensure(ctx->ff_sum, size_t(nimg) * p.nchunk * 4))) return rc;
    if ((rc = ctx->ensure(ctx->ff_base, size_t(nimg) * p.nchunk * 8))) return rc;
    p.blk_off = static_cast<uint32_t*>(ctx->blk_off.p);
    p.tile_sum = static_cast<uint32_t*>(ctx->tile_sum.p);
    p.tile_base = static_cast<uint64_t*>(ctx->tile_base.p);
    p.img_bits = static_cast<uint64_t*>(ctx->img_bits.p);
    p.img_bytes = p.img_bits + nimg;
    p.ustream = static_cast<uint8_t*>(ctx->ustream.p);
    p.ff_sum = static_cast<uint32_t*>(ctx->ff_sum.p);
    p.ff_base = static_cast<uint64_t*>(ctx->ff_base.p);

    (void)jz_launch(k_block_bits, dim3(p.ntile, nimg), dim3(kEntThreads), 0, st, p);
    (void)jz_launch(k_scan_tiles, dim3(nimg), dim3(1024), 0, st, p);
    // grid-stride helpers: enough CTAs to fill the machine, split evenly over the images
    const uint32_t per_img = std::max<uint32_t>(1u, std::min<uint32_t>(p.nchunk, (uint32_t(ctx->num_sms) * 8u + nimg - 1) / nimg));
    (void)jz_launch(k_zero_ustream, dim3(per_img, nimg), dim3(256), 0, st, p);
    (void)jz_launch(k_scatter, dim3(p.ntile, nimg), dim3(kEntThreads), 0, st, p);
    (void)jz_launch(k_ff_count, dim3(per_img, nimg), dim3(kStuffThreads), 0, st, p);
    (void)jz_launch(k_scan_ff, dim3(nimg), dim3(1024), 0, st, p);
    (void)jz_launch(k_stuff_write, dim3(per_img, nimg), dim3(kStuffThreads), 0, st, p);
    ctx->launches += 7;
    JZ_CUDA_TRY(ctx, cudaGetLastError());
    return JPEZYB200_OK;
}

int jpezyb200_read_sizes(jpezyb200_ctx* ctx, const uint64_t* d_values, uint32_t n, uint64_t* h_out, void* stream)
{
    if (!ctx) return JPEZYB200_EINVAL;
    if (!d_values || !h_out || n == 0 || n > 8192u) return ctx->fail(JPEZYB200_EINVAL, "null pointer / bad count");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->h_sizes) JZ_CUDA_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_sizes), 8192 * sizeof(uint64_t)));
    cudaStream_t st = pick_stream(ctx, stream);
    JZ_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_sizes, d_values, size_t(n) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    JZ_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    std::memcpy(h_out, ctx->h_sizes, size_t(n) * sizeof(uint64_t));
    return JPEZYB200_OK;
}

int jpezyb200_transform_fwd_dev(jpezyb200_ctx* ctx, const uint8_t* d_r, const uint8_t* d_g, const uint8_t* d_b, uint32_t W,
                                uint32_t H, uint32_t nimg, int gray, int16_t* d_coefs, void* stream)
{
    int rc = check_geometry(ctx, W, H, nimg);
    if (rc) return rc;
    if (!d_r || !d_g || !d_b || !d_coefs) return ctx->fail(JPEZYB200_EINVAL, "null pointer");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return launch_fwd(ctx, d_r, d_g, d_b, W, H, nimg, gray, d_coefs, pick_stream(ctx, stream));
}

int jpezyb200_entropy_encode_dev(jpezyb200_ctx* ctx, const int16_t* d_coefs, uint32_t W, uint32_t H, uint32_t nimg, int gray,
                                 uint8_t* d_scan, size_t slot_bytes, uint64_t* d_scan_bytes, uint64_t* d_scan_bits, void* stream)
{
    (void)gray;
    int rc = check_geometry(ctx, W, H, nimg);
    if (rc) return rc;
    if (!d_coefs || !d_scan || slot_bytes == 0) return ctx->fail(JPEZYB200_EINVAL, "null pointer / empty slot");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return launch_entropy(ctx, d_coefs, W, H, nimg, d_scan, slot_bytes, d_scan_bytes, d_scan_bits, pick_stream(ctx, stream));
}

int jpezyb200_encode_batch_dev(jpezyb200_ctx* ctx, const uint8_t* d_r, const uint8_t* d_g, const uint8_t* d_b, uint32_t W,
                               uint32_t H, uint32_t nimg, int gray, uint8_t* d_scan, size_t slot_bytes, uint64_t* d_scan_bytes,
                               uint64_t* d_scan_bits, void* stream)
{
    int rc = check_geometry(ctx, W, H, nimg);
    if (rc) return rc;
    if (!d_r || !d_g || !d_b || !d_scan || slot_bytes == 0) return ctx->fail(JPEZYB200_EINVAL, "null pointer / empty slot");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t ncoef = size_t(mcu_units(W)) * mcu_units(H) * 384 * nimg;
    if ((rc = ctx->ensure(ctx->coefs, ncoef * sizeof(int16_t)))) return rc;
    cudaStream_t st = pick_stream(ctx, stream);
    if ((rc = launch_fwd(ctx, d_r, d_g, d_b, W, H, nimg, gray, static_cast<int16_t*>(ctx->coefs.p), st, 0, 0, true))) return rc;
    return launch_entropy(ctx, static_cast<int16_t*>(ctx->coefs.p), W, H, nimg, d_scan, slot_bytes, d_scan_bytes, d_scan_bits, st, true);
}

int jpezyb200_encode(jpezyb200_ctx* ctx, const uint8_t* r, const uint8_t* g, const uint8_t* b, uint32_t W, uint32_t H, int gray,
                     uint8_t* scan_out, size_t scan_cap, size_t* scan_bytes, uint64_t* scan_bits)
{
    int rc = check_geometry(ctx, W, H, 1);
    if (rc) return rc;
    if (!r || !g || !b || !scan_out || !scan_bytes || scan_cap == 0) return ctx->fail(JPEZYB200_EINVAL, "null pointer / empty buffer");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t npx = size_t(W) * H;
    const uint32_t HU = mcu_units(W), VU = mcu_units(H);
    if ((rc = ctx->ensure(ctx->planes_in, 3 * npx))) return rc;
    if ((rc = ctx->ensure(ctx->scan_io, scan_cap))) return rc;
    if ((rc = ctx->ensure(ctx->sizes_io, 64))) return rc;
    if ((rc = ctx->ensure(ctx->coefs, size_t(HU) * VU * 384 * sizeof(int16_t)))) return rc;
    uint8_t* d_in = static_cast<uint8_t*>(ctx->planes_in.p);
    uint64_t* d_sz = static_cast<uint64_t*>(ctx->sizes_io.p);
    int16_t* d_coefs = static_cast<int16_t*>(ctx->coefs.p);
    cudaStream_t st = ctx->stream;
    // bands of MCU rows: the copy of band k+1 (copy stream) overlaps the transform of band k (compute stream)
    HostPipe* hp;
    if ((rc = host_pipe(ctx, &hp))) return rc;
    const uint32_t nband = uint32_t(std::max<size_t>(1, std::min<size_t>(std::min<size_t>(host_bands(), VU), 3 * npx / (size_t(4) << 20))));
    for (uint32_t k = 0; k < nband; ++k) {
        const uint32_t row0 = uint32_t(uint64_t(VU) * k / nband), row1 = uint32_t(uint64_t(VU) * (k + 1) / nband);
        const size_t y0 = std::min<size_t>(H, size_t(row0) * 16), y1 = std::min<size_t>(H, size_t(row1) * 16);
        // (edge replication reads the last image row: the last band carries it)
        const size_t off = y0 * W, len = (y1 - y0) * W;
        if (len) {
            JZ_CUDA_TRY(ctx, cudaMemcpyAsync(d_in + off, r + off, len, cudaMemcpyHostToDevice, hp->copy));
            JZ_CUDA_TRY(ctx, cudaMemcpyAsync(d_in + npx + off, g + off, len, cudaMemcpyHostToDevice, hp->copy));
            JZ_CUDA_TRY(ctx, cudaMemcpyAsync(d_in + 2 * npx + off, b + off, len, cudaMemcpyHostToDevice, hp->copy));
        }
        JZ_CUDA_TRY(ctx, cudaEventRecord(hp->ev[k], hp->copy));
        JZ_CUDA_TRY(ctx, cudaStreamWaitEvent(st, hp->ev[k], 0));
        if ((rc = launch_fwd(ctx, d_in, d_in + npx, d_in + 2 * npx, W, H, 1, gray, d_coefs, st, row0, row1 - row0, true))) return rc;
    }
    if ((rc = launch_entropy(ctx, d_coefs, W, H, 1, static_cast<uint8_t*>(ctx->scan_io.p), scan_cap, d_sz, d_sz + 1, st, true))) return rc;
    uint64_t* h_sz = hp->h_sz;
    JZ_CUDA_TRY(ctx, cudaMemcpyAsync(h_sz, d_sz, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    JZ_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_sz[0] == ~0ull || h_sz[0] > scan_cap) return ctx->fail(JPEZYB200_ECAPACITY, "entropy-coded segment does not fit in scan_cap");
    JZ_CUDA_TRY(ctx, cudaMemcpyAsync(scan_out, ctx->scan_io.p, h_sz[0], cudaMemcpyDeviceToHost, st));
    JZ_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    *scan_bytes = size_t(h_sz[0]);
    if (scan_bits) *scan_bits = h_sz[1];
    return JPEZYB200_OK;
}

// -------------------------------------------------------------------------------------------------
// synthetic input
// -------------------------------------------------------------------------------------------------
int jpezyb200_synth_dev(jpezyb200_ctx* ctx, uint8_t* d_r, uint8_t* d_g, uint8_t* d_b, uint32_t W, uint32_t H, uint32_t nimg,
                        uint32_t first_frame, int family, void* stream)
{
    int rc = check_geometry(ctx, W, H, nimg);
    if (rc) return rc;
    if (!d_r || !d_g || !d_b) return ctx->fail(JPEZYB200_EINVAL, "null pointer");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t npx = size_t(W) * H;
    const uint32_t gx = uint32_t(std::min<size_t>((npx + 255) / 256, size_t(ctx->num_sms) * 16));
    k_synth<<<dim3(gx, nimg), 256, 0, pick_stream(ctx, stream)>>>(d_r, d_g, d_b, W, H, first_frame, family, 0u);
    ++ctx->launches;
    JZ_CUDA_TRY(ctx, cudaGetLastError());
    return JPEZYB200_OK;
}

int jpezyb200_synth_rows_dev(jpezyb200_ctx* ctx, uint8_t* d_r, uint8_t* d_g, uint8_t* d_b, uint32_t W, uint32_t y0, uint32_t nrows,
                             uint32_t frame, int family, void* stream)
{
    int rc = check_geometry(ctx, W, nrows, 1);
    if (rc) return rc;
    if (!d_r || !d_g || !d_b) return ctx->fail(JPEZYB200_EINVAL, "null pointer");
    JZ_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t npx = size_t(W) * nrows;
    const uint32_t gx = uint32_t(std::min<size_t>((npx + 255) / 256, size_t(ctx->num_sms) * 16));
    k_synth<<<dim3(gx, 1), 256, 0, pick_stream(ctx, stream)>>>(d_r, d_g, d_b, W, nrows, frame, family, y0);
    ++ctx->launches;
    JZ_CUDA_TRY(ctx, cudaGetLastError());
    return JPEZYB200_OK;
}

}  // extern "C"

#include "capi_decode.inc"
#include "capi_shard.inc"
#include "capi_batch.inc"
#include "capi_group.inc"
