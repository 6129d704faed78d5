// C++ host, no Python: one image encoded by a group of ranks through jpezyb200_group_encode must give, byte for byte, the
// segment jpezyb200_encode_batch_dev gives for the whole image on one GPU.  Ranks: the devices named on the command line
// (the same device several times = emulated ranks).   usage: group_encode_check W H family gray dev0 [dev1 ...]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/jpezy_b200.h"

#define CHECK(x)                                                                                  \
    do {                                                                                          \
        if (!(x)) {                                                                               \
            std::fprintf(stderr, "FAILED %s (line %d)\n", #x, __LINE__);                          \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

int main(int argc, char** argv)
{
    if (argc < 6) return 2;
    const uint32_t W = uint32_t(std::atoi(argv[1])), H = uint32_t(std::atoi(argv[2]));
    const int family = std::atoi(argv[3]), gray = std::atoi(argv[4]);
    std::vector<int> devs;
    for (int i = 5; i < argc; ++i) devs.push_back(std::atoi(argv[i]));
    const int n = int(devs.size());
    const size_t npx = size_t(W) * H, cap = npx * 3 + 10240;

    // the whole image on device devs[0], encoded by one context
    jpezyb200_ctx* ctx = nullptr;
    CHECK(jpezyb200_ctx_create(devs[0], &ctx) == JPEZYB200_OK);
    CHECK(cudaSetDevice(devs[0]) == cudaSuccess);
    uint8_t *full = nullptr, *one = nullptr;
    uint64_t* d_nb = nullptr;
    CHECK(cudaMalloc(&full, 3 * npx) == cudaSuccess && cudaMalloc(&one, cap) == cudaSuccess && cudaMalloc(&d_nb, 16) == cudaSuccess);
    CHECK(jpezyb200_synth_dev(ctx, full, full + npx, full + 2 * npx, W, H, 1, 0, family, nullptr) == JPEZYB200_OK);
    CHECK(jpezyb200_encode_batch_dev(ctx, full, full + npx, full + 2 * npx, W, H, 1, gray, one, cap, d_nb, nullptr, nullptr) == JPEZYB200_OK);
    uint64_t n_one = 0;
    CHECK(jpezyb200_read_sizes(ctx, d_nb, 1, &n_one, nullptr) == JPEZYB200_OK);
    CHECK(n_one > 0 && n_one <= cap);
    std::vector<uint8_t> want(n_one);
    CHECK(cudaMemcpy(want.data(), one, n_one, cudaMemcpyDeviceToHost) == cudaSuccess);

    // the group: every rank gets the pixel rows of its MCU rows, generated in place on its own device
    jpezyb200_group* g = nullptr;
    CHECK(jpezyb200_group_create(n, devs.data(), &g) == JPEZYB200_OK);
    CHECK(jpezyb200_group_size(g) == n);
    std::vector<const uint8_t*> pr(n), pg(n), pb(n);
    std::vector<uint8_t*> owned;
    for (int k = 0; k < n; ++k) {
        uint32_t row0, rows;
        CHECK(jpezyb200_group_partition(g, H, uint32_t(k), &row0, &rows) == JPEZYB200_OK);
        const uint32_t y0 = row0 * 16, y1 = (row0 + rows) * 16 < H ? (row0 + rows) * 16 : H, ny = y1 - y0;
        CHECK(cudaSetDevice(devs[k]) == cudaSuccess);
        uint8_t* p = nullptr;
        CHECK(cudaMalloc(&p, 3 * size_t(ny) * W) == cudaSuccess);
        owned.push_back(p);
        CHECK(jpezyb200_synth_rows_dev(jpezyb200_group_ctx(g, uint32_t(k)), p, p + size_t(ny) * W, p + 2 * size_t(ny) * W, W, y0, ny, 0, family, nullptr) == JPEZYB200_OK);
        CHECK(cudaDeviceSynchronize() == cudaSuccess);
        pr[k] = p, pg[k] = p + size_t(ny) * W, pb[k] = p + 2 * size_t(ny) * W;
    }
    CHECK(cudaSetDevice(devs[0]) == cudaSuccess);
    uint8_t* dst = nullptr;
    CHECK(cudaMalloc(&dst, cap) == cudaSuccess && cudaMemset(dst, 0xEE, cap) == cudaSuccess);
    uint64_t n_grp = 0;
    for (int rep = 0; rep < 2; ++rep) {          // twice: the group is reusable
        const int rc = jpezyb200_group_encode(g, pr.data(), pg.data(), pb.data(), W, H, gray, dst, cap, &n_grp);
        if (rc != JPEZYB200_OK) std::fprintf(stderr, "group_encode: %s: %s\n", jpezyb200_strerror(rc), jpezyb200_group_last_error(g));
        CHECK(rc == JPEZYB200_OK);
        CHECK(n_grp == n_one);
        std::vector<uint8_t> got(n_grp);
        CHECK(cudaMemcpy(got.data(), dst, n_grp, cudaMemcpyDeviceToHost) == cudaSuccess);
        CHECK(std::memcmp(got.data(), want.data(), n_one) == 0);
    }
    // a destination that is too small must be reported, not overrun
    uint64_t dummy = 0;
    CHECK(jpezyb200_group_encode(g, pr.data(), pg.data(), pb.data(), W, H, gray, dst, 1, &dummy) == JPEZYB200_ECAPACITY);
    std::printf("group_encode ok: %ux%u, %d ranks, %llu bytes, identical to the single-GPU segment\n", W, H, n, (unsigned long long)n_one);
    jpezyb200_group_destroy(g);
    for (uint8_t* p : owned) cudaFree(p);
    cudaFree(dst), cudaFree(full), cudaFree(one), cudaFree(d_nb);
    jpezyb200_ctx_destroy(ctx);
    return 0;
}
