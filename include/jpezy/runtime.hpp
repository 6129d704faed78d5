// include/jpezy/runtime.hpp -- the one jpezyb200 context the host classes share (one per process, device chosen by
// JPEZY_B200_DEVICE, default 0).  There is no CPU path: without a CUDA device every hot-path call throws.
#ifndef JPEZY_B200_RUNTIME_HPP
#define JPEZY_B200_RUNTIME_HPP

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../jpezy_b200.h"

namespace jpezy {
namespace b200 {

struct runtime {
    static jpezyb200_ctx* ctx()
    {
        static runtime r;
        return r.ctx_;
    }
    static std::runtime_error error(const char* where, int code)
    {
        return std::runtime_error(std::string(where) + ": " + jpezyb200_strerror(code) + " (" + jpezyb200_last_error(ctx()) + ")");
    }

private:
    runtime()
    {
        const char* d = std::getenv("JPEZY_B200_DEVICE");
        const int rc = jpezyb200_ctx_create(d ? std::atoi(d) : 0, &ctx_);
        if (rc != JPEZYB200_OK) throw std::runtime_error(std::string("jpezy_b200: ") + jpezyb200_strerror(rc));
    }
    ~runtime() { jpezyb200_ctx_destroy(ctx_); }
    jpezyb200_ctx* ctx_ = nullptr;
};

}  // namespace b200
}  // namespace jpezy
#endif
