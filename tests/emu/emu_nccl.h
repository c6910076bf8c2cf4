// tests/emu/emu_nccl.h -- TEST INFRASTRUCTURE: an in-process stand-in for the four NCCL entry points the library binds at run
// time, so that multi-rank runs (one OS thread per rank, each with its own handle and "device") can be exercised on the CPU.
// ncclAllReduce(sum, double) is a rendezvous: the last rank to arrive sums the send buffers IN RANK ORDER (what makes every
// rank receive bit-identical results, like NCCL's deterministic rings do for a fixed topology) and releases the others.
#pragma once
#include <string.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace emu {
struct nccl_uid_t { char internal[128]; };
struct NcclGroup {
    int world = 0, arrived = 0;
    unsigned long gen = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<const void*> send;
    std::vector<void*> recv;
};
struct NcclComm { NcclGroup* g; int rank; };
inline std::mutex& nccl_mu() { static std::mutex m; return m; }
inline std::map<std::string, NcclGroup*>& nccl_groups() { static std::map<std::string, NcclGroup*> m; return m; }

inline int ncclGetUniqueId(nccl_uid_t* id) {
    static unsigned long counter = 0;
    std::lock_guard<std::mutex> lk(nccl_mu());
    memset(id->internal, 0, sizeof(id->internal));
    snprintf(id->internal, sizeof(id->internal), "emu-nccl-%lu", ++counter);
    return 0;
}
inline int ncclCommInitRank(void** comm, int world, nccl_uid_t id, int rank) {
    if (world < 1 || rank < 0 || rank >= world) return 4;           // ncclInvalidArgument
    std::lock_guard<std::mutex> lk(nccl_mu());
    NcclGroup*& g = nccl_groups()[std::string(id.internal, strnlen(id.internal, sizeof(id.internal)))];
    if (!g) { g = new NcclGroup(); g->world = world; g->send.assign((size_t)world, nullptr); g->recv.assign((size_t)world, nullptr); }
    if (g->world != world) return 4;
    *comm = new NcclComm{g, rank};
    return 0;
}
inline int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t) {
    if (dtype != 8 || op != 0 || !comm) return 4;                   // ncclDouble, ncclSum only
    NcclComm* c = (NcclComm*)comm;
    NcclGroup* g = c->g;
    std::unique_lock<std::mutex> lk(g->mu);
    g->send[(size_t)c->rank] = send;
    g->recv[(size_t)c->rank] = recv;
    const unsigned long gen = g->gen;
    if (++g->arrived == g->world) {
        std::vector<double> sum(count, 0.0);
        for (int r = 0; r < g->world; ++r) {
            const double* s = (const double*)g->send[(size_t)r];
            for (size_t i = 0; i < count; ++i) sum[i] += s[i];
        }
        for (int r = 0; r < g->world; ++r) memcpy(g->recv[(size_t)r], sum.data(), count * sizeof(double));
        g->arrived = 0;
        ++g->gen;
        g->cv.notify_all();
    } else {
        g->cv.wait(lk, [&] { return g->gen != gen; });
    }
    return 0;
}
inline int ncclCommDestroy(void* comm) { delete (NcclComm*)comm; return 0; }
inline const char* ncclGetErrorString(int) { return "emulated NCCL error"; }
}  // namespace emu
