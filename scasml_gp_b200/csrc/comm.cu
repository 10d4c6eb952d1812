// Multi-GPU collective of the path through the C ABI: the ONE all-reduce (sum) of the per-rank partial (u, z) blocks of a
// sample-sharded u_solve / uz_solve, for hosts that do not bring torch.distributed.  NCCL is loaded at run time (dlopen "libnccl.so.2":
// the copy already in the process if the host uses one, e.g. PyTorch's), so the product library has no link-time dependency on it and a
// single-GPU user never touches it.  Reference: the pmap-free data parallelism the north star asks for -- "correction samples are sharded
// across the GPUs of one box, and a single NCCL allreduce over NVLink combines the per-level sums".
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <new>
#include <nccl.h>
#include "common.cuh"

namespace scasml {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        if (api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy) api.handle = h;
    });
    return api.handle ? &api : nullptr;
}

int nccl_check(NcclApi* a, ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return OK;
    set_error(std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
    return ERR_CUDA;
}

}  // namespace
}  // namespace scasml

struct scasml_comm {
    ncclComm_t comm;
    int rank, world;
};

using namespace scasml;

extern "C" {

__attribute__((visibility("default"))) int scasml_comm_unique_id(unsigned char* id128_host) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    SC_REQUIRE(id128_host != nullptr, "comm_unique_id: null pointer");
    NcclApi* a = nccl_api();
    SC_REQUIRE(a != nullptr, "NCCL (libnccl.so.2) not found");
    ncclUniqueId id;
    const int rc = nccl_check(a, a->GetUniqueId(&id), "ncclGetUniqueId");
    if (rc != OK) return rc;
    std::memcpy(id128_host, &id, sizeof(id));
    return OK;
}

__attribute__((visibility("default"))) int scasml_comm_init(const unsigned char* id128_host, int rank, int world, scasml_comm** out) {
    SC_REQUIRE(id128_host && out && world >= 1 && rank >= 0 && rank < world, "comm_init: invalid argument");
    NcclApi* a = nccl_api();
    SC_REQUIRE(a != nullptr, "NCCL (libnccl.so.2) not found");
    ncclUniqueId id;
    std::memcpy(&id, id128_host, sizeof(id));
    scasml_comm* c = new (std::nothrow) scasml_comm();
    SC_REQUIRE(c != nullptr, "comm_init: out of memory");
    c->rank = rank; c->world = world;
    const int rc = nccl_check(a, a->CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");     // the current CUDA device of this process
    if (rc != OK) { delete c; return rc; }
    *out = c;
    return OK;
}

__attribute__((visibility("default"))) int scasml_allreduce_partial(scasml_comm* c, double* buf_dev, long long count, void* stream) {
    SC_REQUIRE(c != nullptr && (buf_dev != nullptr || count == 0) && count >= 0, "allreduce_partial: invalid argument");
    if (count == 0 || c->world == 1) return OK;
    NcclApi* a = nccl_api();
    SC_REQUIRE(a != nullptr, "NCCL (libnccl.so.2) not found");
    return nccl_check(a, a->AllReduce(buf_dev, buf_dev, (size_t)count, ncclDouble, ncclSum, c->comm, (cudaStream_t)stream), "ncclAllReduce");
}

__attribute__((visibility("default"))) int scasml_comm_destroy(scasml_comm* c) {
    if (c == nullptr) return OK;
    NcclApi* a = nccl_api();
    int rc = OK;
    if (a != nullptr) rc = nccl_check(a, a->CommDestroy(c->comm), "ncclCommDestroy");
    delete c;
    return rc;
}

}  // extern "C"
