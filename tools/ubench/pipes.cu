// pipes.cu -- instruction-throughput probes for sm_100a (B200).  Not part of the product: it answers the design questions
// of the transform kernels (which conversions avoid the XU pipe, what packed f32x2 / IDP / PRMT / 3-input min cost, how the
// FMA and ALU pipes co-issue).  Each probe runs U independent dependency chains per thread, 8 warps per scheduler, one CTA
// per SM, and reports warp-instructions per clock per SM (4.0 = every scheduler issues each cycle).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench_pipes tools/ubench/pipes.cu && build/ubench_pipes
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

constexpr int kIters = 512;
constexpr int U = 8;

#define PROBE_KERNEL(NAME, DECL, BODY, SINK)                                                            \
    __global__ void __launch_bounds__(1024, 1) NAME(uint32_t* out, long long* cyc, uint32_t seed)        \
    {                                                                                                    \
        DECL;                                                                                            \
        __syncthreads();                                                                                 \
        const long long t0 = clock64();                                                                  \
        _Pragma("unroll 1") for (int it = 0; it < kIters; ++it) { BODY; }                               \
        const long long t1 = clock64();                                                                  \
        __syncthreads();                                                                                 \
        uint32_t s = 0;                                                                                  \
        SINK;                                                                                            \
        if (s == 0x12345678u) out[threadIdx.x] = s;                                                      \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                 \
    }

// ---------------- 32-bit register chains ----------------
#define DECL_R uint32_t r[U]; _Pragma("unroll") for (int i = 0; i < U; ++i) r[i] = seed * (threadIdx.x + 1) + i * 77u; uint32_t k0 = seed | 1u, k1 = seed ^ 0x5140u
#define SINK_R _Pragma("unroll") for (int i = 0; i < U; ++i) s ^= r[i]
#define DECL_F float f[U]; _Pragma("unroll") for (int i = 0; i < U; ++i) f[i] = float(seed * (threadIdx.x + 1) + i) * 1e-3f; float c0 = float(seed) * 1e-4f + 1.0f, c1 = float(seed) * 1e-5f
#define SINK_F _Pragma("unroll") for (int i = 0; i < U; ++i) s ^= __float_as_uint(f[i])
#define DECL_L unsigned long long l[U]; _Pragma("unroll") for (int i = 0; i < U; ++i) l[i] = (unsigned long long)(__float_as_uint(float(seed + i) * 1e-3f)) * 0x100000001ull; unsigned long long lc = (unsigned long long)__float_as_uint(1.0001f) * 0x100000001ull
#define SINK_L _Pragma("unroll") for (int i = 0; i < U; ++i) s ^= uint32_t(l[i]) ^ uint32_t(l[i] >> 32)
#define DECL_D double d[U]; _Pragma("unroll") for (int i = 0; i < U; ++i) d[i] = double(seed + i) * 1e-3; double dc = double(seed) * 1e-6 + 1.0
#define SINK_D _Pragma("unroll") for (int i = 0; i < U; ++i) s ^= uint32_t(__double_as_longlong(d[i]))

#define REP(X) _Pragma("unroll") for (int i = 0; i < U; ++i) { X; }

PROBE_KERNEL(k_ffma, DECL_F, REP(asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c0), "f"(c1))), SINK_F)
PROBE_KERNEL(k_ffma_imm, DECL_F, REP(asm volatile("fma.rn.f32 %0, %0, 0f3F800347, 0f3A83126F;" : "+f"(f[i]))), SINK_F)
PROBE_KERNEL(k_fadd, DECL_F, REP(asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1))), SINK_F)
PROBE_KERNEL(k_fadd_rz, DECL_F, REP(asm volatile("add.rz.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1))), SINK_F)
PROBE_KERNEL(k_fmul, DECL_F, REP(asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c0))), SINK_F)
PROBE_KERNEL(k_ffma2, DECL_L, REP(asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(l[i]) : "l"(lc))), SINK_L)
PROBE_KERNEL(k_ffma2_rz, DECL_L, REP(asm volatile("fma.rz.f32x2 %0, %0, %1, %1;" : "+l"(l[i]) : "l"(lc))), SINK_L)
PROBE_KERNEL(k_fadd2, DECL_L, REP(asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(l[i]) : "l"(lc))), SINK_L)
PROBE_KERNEL(k_fadd2_rm, DECL_L, REP(asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(l[i]) : "l"(lc))), SINK_L)
PROBE_KERNEL(k_fmul2, DECL_L, REP(asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(l[i]) : "l"(lc))), SINK_L)
PROBE_KERNEL(k_prmt, DECL_R, REP(asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)
PROBE_KERNEL(k_lop3, DECL_R, REP(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)
PROBE_KERNEL(k_iadd3, DECL_R, REP(asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(k0))), SINK_R)
PROBE_KERNEL(k_shf, DECL_R, REP(asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(k0))), SINK_R)
PROBE_KERNEL(k_imad, DECL_R, REP(asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)
PROBE_KERNEL(k_imadhi, DECL_R, REP(asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(k0))), SINK_R)
PROBE_KERNEL(k_dp2a, DECL_R, REP(asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)
PROBE_KERNEL(k_dp4a, DECL_R, REP(asm volatile("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)
PROBE_KERNEL(k_i2f, DECL_R, REP(asm volatile("{.reg .f32 t; cvt.rn.f32.s32 t, %0; mov.b32 %0, t;}" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_i2f_u8, DECL_R, REP(asm volatile("{.reg .f32 t; .reg .b16 lo, hi; mov.b32 {lo, hi}, %0; cvt.rn.f32.u8 t, lo; mov.b32 %0, t;}" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_f2i_rz, DECL_F, REP(asm volatile("{.reg .s32 t; cvt.rzi.s32.f32 t, %0; mov.b32 %0, t;}" : "+f"(f[i]))), SINK_F)
PROBE_KERNEL(k_frnd_rz, DECL_F, REP(asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(f[i]))), SINK_F)
PROBE_KERNEL(k_fmnmx, DECL_F, REP(asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1))), SINK_F)
PROBE_KERNEL(k_fmnmx3, DECL_F, REP(asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c0))), SINK_F)
PROBE_KERNEL(k_fmnmx_abs, DECL_F, REP(asm volatile("{.reg .f32 t; abs.f32 t, %0; max.f32 %0, t, %1;}" : "+f"(f[i]) : "f"(c1))), SINK_F)
PROBE_KERNEL(k_fsetp_sel, DECL_F, REP(asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(f[i]) : "f"(c1), "f"(c0))), SINK_F)
PROBE_KERNEL(k_i2ip, DECL_R, REP(asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)
PROBE_KERNEL(k_f2fp, DECL_R, REP(asm volatile("{.reg .f32 a; mov.b32 a, %0; cvt.rn.f16x2.f32 %0, a, a;}" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_hadd2, DECL_R, REP(asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(k0))), SINK_R)
PROBE_KERNEL(k_hfma2, DECL_R, REP(asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)
PROBE_KERNEL(k_hmnmx2, DECL_R, REP(asm volatile("max.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(k0))), SINK_R)
PROBE_KERNEL(k_dadd, DECL_D, REP(asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dc))), SINK_D)
PROBE_KERNEL(k_dfma, DECL_D, REP(asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(dc))), SINK_D)
PROBE_KERNEL(k_shfl, DECL_R, REP(asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_vote, DECL_R, REP(asm volatile("{.reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_redux, DECL_R, REP(asm volatile("redux.sync.or.b32 %0, %0, 0xffffffff;" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_popc, DECL_R, REP(asm volatile("popc.b32 %0, %0;" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_flo, DECL_R, REP(asm volatile("clz.b32 %0, %0;" : "+r"(r[i]))), SINK_R)
PROBE_KERNEL(k_vabsdiff, DECL_R, REP(asm volatile("sad.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k0), "r"(k1))), SINK_R)

// mixes: one FMA-pipe instruction + one ALU-pipe instruction per chain step (2 instructions counted)
PROBE_KERNEL(k_mix_ffma_lop3, DECL_F; uint32_t r[U]; REP(r[i] = seed + i), REP(asm volatile("fma.rn.f32 %0, %0, %2, %3; lop3.b32 %1, %1, %4, %4, 0x96;" : "+f"(f[i]), "+r"(r[i]) : "f"(c0), "f"(c1), "r"(seed))), SINK_F; REP(s ^= r[i]))
PROBE_KERNEL(k_mix_ffma_prmt, DECL_F; uint32_t r[U]; REP(r[i] = seed + i), REP(asm volatile("fma.rn.f32 %0, %0, %2, %3; prmt.b32 %1, %1, %4, %4;" : "+f"(f[i]), "+r"(r[i]) : "f"(c0), "f"(c1), "r"(seed))), SINK_F; REP(s ^= r[i]))
PROBE_KERNEL(k_mix_ffma2_prmt, DECL_L; uint32_t r[U]; REP(r[i] = seed + i), REP(asm volatile("fma.rn.f32x2 %0, %0, %2, %2; prmt.b32 %1, %1, %3, %3;" : "+l"(l[i]), "+r"(r[i]) : "l"(lc), "r"(seed))), SINK_L; REP(s ^= r[i]))
PROBE_KERNEL(k_mix_ffma2_2prmt, DECL_L; uint32_t r[U]; REP(r[i] = seed + i), REP(asm volatile("fma.rn.f32x2 %0, %0, %2, %2; prmt.b32 %1, %1, %3, %3; lop3.b32 %1, %1, %3, %3, 0x96;" : "+l"(l[i]), "+r"(r[i]) : "l"(lc), "r"(seed))), SINK_L; REP(s ^= r[i]))
PROBE_KERNEL(k_mix_ffma_dp2a, DECL_F; uint32_t r[U]; REP(r[i] = seed + i), REP(asm volatile("fma.rn.f32 %0, %0, %2, %3; dp2a.lo.s32.u32 %1, %4, %4, %1;" : "+f"(f[i]), "+r"(r[i]) : "f"(c0), "f"(c1), "r"(seed))), SINK_F; REP(s ^= r[i]))
PROBE_KERNEL(k_mix_prmt_dp2a, DECL_R; uint32_t q[U]; REP(q[i] = seed + i), REP(asm volatile("prmt.b32 %0, %0, %2, %2; dp2a.lo.s32.u32 %1, %2, %2, %1;" : "+r"(r[i]), "+r"(q[i]) : "r"(seed))), SINK_R; REP(s ^= q[i]))
PROBE_KERNEL(k_mix_ffma2_fadd, DECL_L; float f[U]; REP(f[i] = float(seed + i)), REP(asm volatile("fma.rn.f32x2 %0, %0, %2, %2; add.rn.f32 %1, %1, %3;" : "+l"(l[i]), "+f"(f[i]) : "l"(lc), "f"(1.5f))), SINK_L; REP(s ^= __float_as_uint(f[i])))
PROBE_KERNEL(k_mix_ffma_i2f, DECL_F; uint32_t r[U]; REP(r[i] = seed + i), REP(asm volatile("{.reg .f32 t; fma.rn.f32 %0, %0, %2, %3; fma.rn.f32 %0, %0, %2, %3; fma.rn.f32 %0, %0, %2, %3; cvt.rn.f32.s32 t, %1; mov.b32 %1, t;}" : "+f"(f[i]), "+r"(r[i]) : "f"(c0), "f"(c1))), SINK_F; REP(s ^= r[i]))

// shared memory
__global__ void __launch_bounds__(1024, 1) k_lds64(uint32_t* out, long long* cyc, uint32_t seed)
{
    __shared__ __align__(16) unsigned long long sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 1024) sm[i] = seed + i;
    __syncthreads();
    unsigned long long acc = 0;
    uint32_t a = threadIdx.x & 1023;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) acc ^= sm[(a + i * 32 + it) & 4095];
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[threadIdx.x] = uint32_t(acc);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024, 1) k_lds128(uint32_t* out, long long* cyc, uint32_t seed)
{
    __shared__ __align__(16) uint4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 1024) sm[i] = make_uint4(seed, i, 0, 1);
    __syncthreads();
    uint32_t acc = 0;
    uint32_t a = threadIdx.x & 1023;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const uint4 v = sm[(a + i * 32 + it) & 2047];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024, 1) k_lds32(uint32_t* out, long long* cyc, uint32_t seed)
{
    __shared__ uint32_t sm[8192];
    for (int i = threadIdx.x; i < 8192; i += 1024) sm[i] = seed + i;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t a = threadIdx.x & 1023;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) acc ^= sm[(a + i * 32 + it) & 8191];
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024, 1) k_sts16(uint32_t* out, long long* cyc, uint32_t seed)
{
    __shared__ uint16_t sm[16384];
    uint32_t a = threadIdx.x & 1023;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) sm[(a * 2 + i * 2048 + it) & 16383] = uint16_t(seed + it);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (sm[a] == 0x1234u && seed == 77) out[threadIdx.x] = sm[a];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024, 1) k_sts128(uint32_t* out, long long* cyc, uint32_t seed)
{
    __shared__ __align__(16) uint4 sm[2048];
    uint32_t a = threadIdx.x & 1023;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) sm[(a + i * 32 + it) & 2047] = make_uint4(seed, it, i, a);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (sm[a].x == 0x1234u && seed == 77) out[threadIdx.x] = sm[a].y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Probe {
    const char* name;
    void (*fn)(uint32_t*, long long*, uint32_t);
    int insts_per_step;   // counted instructions per chain step
};

int main(int argc, char** argv)
{
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, dev);
    const int nsm = pr.multiProcessorCount;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 4096 * 4);
    cudaMalloc(&cyc, nsm * 8);
    std::vector<long long> h(nsm);
    const Probe probes[] = {
        {"FFMA (3 reg)", k_ffma, 1},        {"FFMA (imm)", k_ffma_imm, 1},     {"FADD", k_fadd, 1},          {"FADD.RZ", k_fadd_rz, 1},
        {"FMUL", k_fmul, 1},                {"FFMA2", k_ffma2, 1},             {"FFMA2.RZ", k_ffma2_rz, 1},  {"FADD2", k_fadd2, 1},
        {"FADD2.RM", k_fadd2_rm, 1},        {"FMUL2", k_fmul2, 1},             {"PRMT", k_prmt, 1},          {"LOP3", k_lop3, 1},
        {"IADD", k_iadd3, 1},               {"SHF", k_shf, 1},                 {"IMAD", k_imad, 1},          {"IMAD.HI", k_imadhi, 1},
        {"IDP.2A", k_dp2a, 1},              {"IDP.4A", k_dp4a, 1},             {"I2F.S32", k_i2f, 1},        {"I2F.U8", k_i2f_u8, 1},
        {"F2I.RZ", k_f2i_rz, 1},            {"FRND.RZ", k_frnd_rz, 1},         {"FMNMX", k_fmnmx, 1},        {"FMNMX3", k_fmnmx3, 1},
        {"FMNMX |a|", k_fmnmx_abs, 1},      {"FSETP+SEL", k_fsetp_sel, 2},     {"I2IP.SAT", k_i2ip, 1},      {"F2FP.F16x2", k_f2fp, 1},
        {"HADD2", k_hadd2, 1},              {"HFMA2", k_hfma2, 1},             {"HMNMX2", k_hmnmx2, 1},      {"DADD", k_dadd, 1},
        {"DFMA", k_dfma, 1},                {"SHFL", k_shfl, 1},               {"VOTE(+setp)", k_vote, 2},   {"REDUX.OR", k_redux, 1},
        {"POPC", k_popc, 1},                {"FLO/CLZ", k_flo, 1},             {"VABSDIFF", k_vabsdiff, 1},
        {"mix FFMA+LOP3", k_mix_ffma_lop3, 2},   {"mix FFMA+PRMT", k_mix_ffma_prmt, 2},   {"mix FFMA2+PRMT", k_mix_ffma2_prmt, 2},
        {"mix FFMA2+PRMT+LOP3", k_mix_ffma2_2prmt, 3}, {"mix FFMA+IDP.2A", k_mix_ffma_dp2a, 2}, {"mix PRMT+IDP.2A", k_mix_prmt_dp2a, 2},
        {"mix FFMA2+FADD", k_mix_ffma2_fadd, 2}, {"mix 3FFMA+I2F", k_mix_ffma_i2f, 4},
        {"LDS.64", k_lds64, 1},             {"LDS.128", k_lds128, 1},          {"LDS.32", k_lds32, 1},       {"STS.U16", k_sts16, 1},
        {"STS.128", k_sts128, 1},
    };
    printf("%s, %d SMs, clock %d kHz\n", pr.name, nsm, pr.clockRate);
    printf("%-24s %12s %14s\n", "probe", "cycles", "warp-inst/clk/SM");
    for (const Probe& p : probes) {
        for (int rep = 0; rep < 2; ++rep) {
            p.fn<<<nsm, 1024>>>(out, cyc, 3u);
            cudaDeviceSynchronize();
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            printf("%-24s error %s\n", p.name, cudaGetErrorString(e));
            continue;
        }
        cudaMemcpy(h.data(), cyc, nsm * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < nsm; ++i) avg += double(h[i]);
        avg /= nsm;
        const double insts = double(kIters) * U * p.insts_per_step * 32.0;   // per SM: 32 warps
        printf("%-24s %12.0f %14.3f\n", p.name, avg, insts / avg);
    }
    return 0;
}
