// az_comm_nccl.cpp -- AzComm over NCCL (NVLink 5 / NVSwitch inside one box).  libnccl is loaded
// with dlopen at the first multi-GPU create, so the single-GPU library has no NCCL dependency and
// a process that already loaded NCCL (e.g. through torch) shares that copy.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>
// No <nccl.h>: the library is loaded with dlopen, and a CUDA-only box (no NCCL headers) must still build the engine for
// single-GPU use.  The handful of NCCL types this file needs are declared here with the values of nccl.h (stable since
// NCCL 2.0: ncclSuccess = 0, ncclChar = 0, ncclDouble = 8, ncclSum = 0, ncclMax = 2, a 128-byte unique id).
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclChar = 0, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2 } ncclRedOp_t;
}

#include "az_comm.h"
#include "../../include/aztot_b200.h"

namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_api;

bool load_api(std::string* err)
{
    if (g_api.h) return true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { if (err) *err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define AZ_SYM(field, name) *(void**)(&g_api.field) = dlsym(h, name); if (!g_api.field) { if (err) *err = std::string("libnccl lacks ") + name; return false; }
    AZ_SYM(GetUniqueId, "ncclGetUniqueId") AZ_SYM(CommInitRank, "ncclCommInitRank") AZ_SYM(CommDestroy, "ncclCommDestroy")
    AZ_SYM(Send, "ncclSend") AZ_SYM(Recv, "ncclRecv") AZ_SYM(GroupStart, "ncclGroupStart") AZ_SYM(GroupEnd, "ncclGroupEnd")
    AZ_SYM(AllReduce, "ncclAllReduce") AZ_SYM(AllGather, "ncclAllGather") AZ_SYM(GetErrorString, "ncclGetErrorString")
#undef AZ_SYM
    g_api.h = h;
    return true;
}

struct NcclComm : AzComm {
    ncclComm_t comm = nullptr;
    std::vector<void*> opened; // cudaIpcOpenMemHandle mappings of the peers' buffers
    ~NcclComm() override
    {
        for (void* p : opened) cudaIpcCloseMemHandle(p);
        if (comm) g_api.CommDestroy(comm);
    }
    // Peer memory through CUDA IPC: every rank publishes (handle of the allocation, offset of the buffer inside it) for each
    // buffer, the records travel in one ncclAllGather, and every rank maps the allocations of the others (peer access is
    // enabled lazily by cudaIpcOpenMemHandle).  One box only: all ranks must be able to reach each other over NVLink / PCIe.
    int share(void* const* local, int nbuf, std::vector<void*>* remote) override
    {
        // pid / dev / ptr: ranks that live in ONE process (threads, as in the tests) cannot open each other's IPC handles; they
        // enable direct peer access between their devices and use the pointer itself (unified virtual addressing)
        struct Rec { cudaIpcMemHandle_t h; unsigned long long off; long long pid; int dev; int pad; unsigned long long ptr; };
        typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
        GetRange get_range = nullptr;
        {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qr;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { err = "cuMemGetAddressRange is not available"; return AZ_ERR_UNSUPPORTED; }
            get_range = (GetRange)fn;
        }
        std::vector<Rec> mine(nbuf), all((size_t)nbuf * nranks);
        for (int b = 0; b < nbuf; b++) {
            unsigned long long base = 0; size_t size = 0;
            if (get_range(&base, &size, (unsigned long long)local[b]) != 0) { err = "cuMemGetAddressRange failed"; return AZ_ERR_UNSUPPORTED; }
            if (cudaIpcGetMemHandle(&mine[b].h, (void*)base) != cudaSuccess) { cudaGetLastError(); err = "cudaIpcGetMemHandle failed (peer memory needs cudaMalloc'ed buffers)"; return AZ_ERR_UNSUPPORTED; }
            mine[b].off = (unsigned long long)local[b] - base;
            mine[b].pid = (long long)getpid(); mine[b].ptr = (unsigned long long)local[b]; mine[b].pad = 0;
            cudaGetDevice(&mine[b].dev);
        }
        Rec *d_mine = nullptr, *d_all = nullptr;
        if (cudaMalloc(&d_mine, sizeof(Rec) * nbuf) != cudaSuccess || cudaMalloc(&d_all, sizeof(Rec) * nbuf * nranks) != cudaSuccess) { err = "cudaMalloc failed"; return AZ_ERR_CUDA; }
        cudaMemcpy(d_mine, mine.data(), sizeof(Rec) * nbuf, cudaMemcpyHostToDevice);
        ncclResult_t r = g_api.AllGather(d_mine, d_all, sizeof(Rec) * nbuf, ncclChar, comm, 0);
        if (r != ncclSuccess) { cudaFree(d_mine); cudaFree(d_all); return fail("ncclAllGather", r); }
        cudaStreamSynchronize(0);
        cudaMemcpy(all.data(), d_all, sizeof(Rec) * nbuf * nranks, cudaMemcpyDeviceToHost);
        cudaFree(d_mine); cudaFree(d_all);
        remote->assign((size_t)nbuf * nranks, nullptr);
        int ok = 1;
        for (int rk = 0; rk < nranks && ok; rk++)
            for (int b = 0; b < nbuf; b++) {
                if (rk == rank) { (*remote)[(size_t)rk * nbuf + b] = local[b]; continue; }
                const Rec& rr = all[(size_t)rk * nbuf + b];
                if (rr.pid == (long long)getpid()) { // same process: peer access instead of IPC
                    int mydev = 0, can = 0;
                    cudaGetDevice(&mydev);
                    if (rr.dev != mydev) {
                        cudaDeviceCanAccessPeer(&can, mydev, rr.dev);
                        if (!can) { ok = 0; break; }
                        cudaError_t pe = cudaDeviceEnablePeerAccess(rr.dev, 0);
                        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0; break; }
                        cudaGetLastError();
                    }
                    (*remote)[(size_t)rk * nbuf + b] = (void*)rr.ptr;
                    continue;
                }
                void* p = nullptr;
                // buffers of one rank may share an allocation: map every distinct handle once
                for (int q = 0; q < b; q++)
                    if (!memcmp(&all[(size_t)rk * nbuf + q].h, &all[(size_t)rk * nbuf + b].h, sizeof(cudaIpcMemHandle_t)))
                        p = (char*)(*remote)[(size_t)rk * nbuf + q] - all[(size_t)rk * nbuf + q].off;
                if (!p) {
                    if (cudaIpcOpenMemHandle(&p, all[(size_t)rk * nbuf + b].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
                    opened.push_back(p);
                }
                (*remote)[(size_t)rk * nbuf + b] = (char*)p + all[(size_t)rk * nbuf + b].off;
            }
        // collective verdict: peer memory is used by all ranks or by none
        double v = ok ? 0.0 : 1.0, *d_v = nullptr;
        cudaMalloc(&d_v, sizeof(double));
        cudaMemcpy(d_v, &v, sizeof(double), cudaMemcpyHostToDevice);
        r = g_api.AllReduce(d_v, d_v, 1, ncclDouble, ncclMax, comm, 0);
        cudaStreamSynchronize(0);
        cudaMemcpy(&v, d_v, sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(d_v);
        if (r != ncclSuccess) return fail("ncclAllReduce", r);
        if (v != 0.0) { err = "cudaIpcOpenMemHandle failed on some rank (no peer access between the GPUs?)"; return AZ_ERR_UNSUPPORTED; }
        return AZ_OK;
    }
    int fail(const char* what, ncclResult_t r)
    {
        err = std::string(what) + ": " + g_api.GetErrorString(r);
        return AZ_ERR_COMM;
    }
    int exchange(const std::vector<AzP2P>& ops, cudaStream_t st) override
    {
        ncclResult_t r = g_api.GroupStart();
        if (r != ncclSuccess) return fail("ncclGroupStart", r);
        for (const AzP2P& o : ops)
            if (o.send_bytes) { r = g_api.Send(o.send, o.send_bytes, ncclChar, o.peer, comm, st); if (r != ncclSuccess) { g_api.GroupEnd(); return fail("ncclSend", r); } }
        for (const AzP2P& o : ops)
            if (o.recv_bytes) { r = g_api.Recv(o.recv, o.recv_bytes, ncclChar, o.peer, comm, st); if (r != ncclSuccess) { g_api.GroupEnd(); return fail("ncclRecv", r); } }
        r = g_api.GroupEnd();
        if (r != ncclSuccess) return fail("ncclGroupEnd", r);
        return AZ_OK;
    }
    int allreduce(double* dev, int n, bool take_max, cudaStream_t st) override
    {
        ncclResult_t r = g_api.AllReduce(dev, dev, (size_t)n, ncclDouble, take_max ? ncclMax : ncclSum, comm, st);
        if (r != ncclSuccess) return fail("ncclAllReduce", r);
        return AZ_OK;
    }
};
} // namespace

int az_comm_make_id(void* id128, std::string* err)
{
    if (!load_api(err)) return AZ_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_api.GetUniqueId(&id);
    if (r != ncclSuccess) { if (err) *err = std::string("ncclGetUniqueId: ") + g_api.GetErrorString(r); return AZ_ERR_COMM; }
    memcpy(id128, &id, 128);
    return AZ_OK;
}

int az_comm_create(const void* id, int rank, int nranks, int device, AzComm** out, std::string* err)
{
    if (!id) { if (err) *err = "multi-GPU create needs opts.nccl_id (128 bytes from az_comm_unique_id, the same on all ranks)"; return AZ_ERR_INVALID; }
    if (!load_api(err)) return AZ_ERR_COMM;
    if (cudaSetDevice(device) != cudaSuccess) { if (err) *err = "cudaSetDevice failed"; return AZ_ERR_CUDA; }
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    NcclComm* c = new NcclComm();
    c->rank = rank; c->nranks = nranks;
    ncclResult_t r = g_api.CommInitRank(&c->comm, nranks, uid, rank);
    if (r != ncclSuccess) { if (err) *err = std::string("ncclCommInitRank: ") + g_api.GetErrorString(r); delete c; return AZ_ERR_COMM; }
    *out = c;
    return AZ_OK;
}
