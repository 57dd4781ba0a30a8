// sfk_comm.cpp -- NCCL, bound at run time.
//
// The one exchange step of the path -- a halo whose particles are spread over several GPUs,
// the analogue of Halo.Split / Halo.Join (los/halo.go:107-140) -- runs over NCCL inside the
// library (sfk_api.cu: sfk_comm_init, sfk_split_finish_comm).  libnccl is opened with dlopen
// on first use, so single-GPU callers need no NCCL at all and a host process that already
// holds an NCCL (PyTorch's bundled copy) shares it instead of loading a second one.
// Search order: a libnccl.so.2 already mapped into the process, $SFK_NCCL_LIB, then the
// loader's default path.
#include <dlfcn.h>

#include <mutex>
#include <string>

#include "sfk_comm.h"

namespace sfk {

namespace {
NcclApi g_api;
bool g_loaded = false;
std::string g_err;
std::mutex g_mu;

template <typename F>
bool bind(void *h, const char *name, F &fn) {
    fn = reinterpret_cast<F>(dlsym(h, name));
    if (!fn) g_err = std::string("libnccl: missing symbol ") + name;
    return fn != nullptr;
}
}  // namespace

const NcclApi *nccl_api(std::string *err) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_loaded && g_err.empty()) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) {
            const char *env = getenv("SFK_NCCL_LIB");
            if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            g_err = std::string("cannot load libnccl.so.2 (set SFK_NCCL_LIB): ") + dlerror();
        } else if (bind(h, "ncclGetUniqueId", g_api.GetUniqueId) && bind(h, "ncclCommInitRank", g_api.CommInitRank) &&
                   bind(h, "ncclCommDestroy", g_api.CommDestroy) && bind(h, "ncclAllReduce", g_api.AllReduce) &&
                   bind(h, "ncclReduceScatter", g_api.ReduceScatter) && bind(h, "ncclAllGather", g_api.AllGather) &&
                   bind(h, "ncclGroupStart", g_api.GroupStart) && bind(h, "ncclGroupEnd", g_api.GroupEnd) &&
                   bind(h, "ncclGetErrorString", g_api.GetErrorString) && bind(h, "ncclGetVersion", g_api.GetVersion)) {
            g_loaded = true;
        }
    }
    if (!g_loaded) {
        if (err) *err = g_err;
        return nullptr;
    }
    return &g_api;
}

}  // namespace sfk
