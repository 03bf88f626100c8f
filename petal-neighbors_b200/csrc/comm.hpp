// comm.hpp -- NCCL inside the library: one pn_comm is one rank of a communicator bound to one device.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2) instead of being a link-time dependency: a process that has
// already loaded a NCCL (PyTorch brings its own) shares that copy, a C++ / Rust host gets the system one, and a
// single-GPU user of the library needs no NCCL at all.  Only the declarations of <nccl.h> are used at compile time.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

namespace petal {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;

    // the process-wide table; nullptr (with *err set) when no NCCL can be loaded
    static NcclApi* get(std::string* err) {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            const char* env = getenv("PN_NCCL_LIB");
            const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
            for (const char* nm : names) {
                if (!nm || !*nm) continue;
                api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
                if (api.lib) break;
            }
            if (!api.lib) { api.error = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror(); return; }
            bool ok = true;
            auto sym = [&](auto& fn, const char* name) {
                fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.lib, name));
                if (!fn) { ok = false; api.error = std::string("NCCL symbol missing: ") + name; }
            };
            sym(api.GetUniqueId, "ncclGetUniqueId"); sym(api.CommInitRank, "ncclCommInitRank"); sym(api.CommInitAll, "ncclCommInitAll");
            sym(api.CommDestroy, "ncclCommDestroy"); sym(api.AllGather, "ncclAllGather"); sym(api.Broadcast, "ncclBroadcast");
            sym(api.Send, "ncclSend"); sym(api.Recv, "ncclRecv"); sym(api.GroupStart, "ncclGroupStart"); sym(api.GroupEnd, "ncclGroupEnd");
            sym(api.GetErrorString, "ncclGetErrorString"); sym(api.GetVersion, "ncclGetVersion");
            if (!ok) { dlclose(api.lib); api.lib = nullptr; }
        });
        if (!api.lib) { if (err) *err = api.error; return nullptr; }
        return &api;
    }
};

}  // namespace petal

// one rank of a communicator (opaque at the C ABI)
struct pn_comm {
    petal::NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;       // the exchange stream: collectives run here, next to the tree's compute stream
    unsigned long long bytes_sent = 0;   // payload bytes this rank handed to NCCL for other ranks (cumulative)
    unsigned long long collectives = 0;  // NCCL calls issued (cumulative)
};
