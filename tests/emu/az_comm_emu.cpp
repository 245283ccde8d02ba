// az_comm_emu.cpp -- TEST INFRASTRUCTURE: AzComm as an in-process mailbox, so that the domain
// decomposition logic can run as N engines on N host threads of one process (tests/emu only).
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include "az_comm.h"
#include "../../include/aztot_b200.h"

namespace {
struct Msg { std::vector<char> data; };
struct World {
    std::mutex m; std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<Msg>> box; // (src,dst) -> FIFO
    int nranks = 0;
    // allreduce state
    int ar_arrived = 0; unsigned ar_gen = 0; std::vector<double> ar_sum, ar_out;
};
std::mutex g_m;
std::map<std::string, World*> g_worlds;

struct EmuComm : AzComm {
    World* w = nullptr;
    int exchange(const std::vector<AzP2P>& ops, cudaStream_t) override
    {
        {
            std::lock_guard<std::mutex> lk(w->m);
            for (const AzP2P& o : ops)
                if (o.send_bytes) { Msg msg; msg.data.assign((const char*)o.send, (const char*)o.send + o.send_bytes); w->box[{ rank, o.peer }].push_back(std::move(msg)); }
        }
        w->cv.notify_all();
        for (const AzP2P& o : ops)
            if (o.recv_bytes) {
                std::unique_lock<std::mutex> lk(w->m);
                auto& q = w->box[{ o.peer, rank }];
                if (!w->cv.wait_for(lk, std::chrono::seconds(60), [&] { return !q.empty(); })) { err = "emu comm: receive timed out"; return AZ_ERR_COMM; }
                Msg msg = std::move(q.front()); q.pop_front();
                if (msg.data.size() != o.recv_bytes) { err = "emu comm: message size mismatch"; return AZ_ERR_COMM; }
                memcpy(o.recv, msg.data.data(), o.recv_bytes);
            }
        return AZ_OK;
    }
    int share(void* const* local, int nbuf, std::vector<void*>* remote) override
    {
        if (getenv("AZ_EMU_NO_PEER")) return AZ_ERR_UNSUPPORTED; // tests: the send/recv halo path
        std::vector<AzP2P> ops;
        std::vector<std::vector<void*>> in(nranks, std::vector<void*>(nbuf));
        for (int r = 0; r < nranks; r++) if (r != rank) ops.push_back({ r, local, sizeof(void*) * nbuf, nullptr, 0 });
        for (int r = 0; r < nranks; r++) if (r != rank) ops.push_back({ r, nullptr, 0, in[r].data(), sizeof(void*) * nbuf });
        int rc = exchange(ops, nullptr);
        if (rc) return rc;
        remote->assign((size_t)nbuf * nranks, nullptr);
        for (int r = 0; r < nranks; r++)
            for (int b = 0; b < nbuf; b++) (*remote)[(size_t)r * nbuf + b] = r == rank ? local[b] : in[r][b];
        return AZ_OK;
    }
    int allreduce(double* dev, int n, bool take_max, cudaStream_t) override
    {
        std::unique_lock<std::mutex> lk(w->m);
        unsigned g = w->ar_gen;
        if (w->ar_arrived == 0) w->ar_sum.assign(dev, dev + n);
        else for (int k = 0; k < n; k++) w->ar_sum[k] = take_max ? std::max(w->ar_sum[k], dev[k]) : w->ar_sum[k] + dev[k];
        if (++w->ar_arrived == w->nranks) { w->ar_out = w->ar_sum; w->ar_arrived = 0; w->ar_gen++; w->cv.notify_all(); }
        else if (!w->cv.wait_for(lk, std::chrono::seconds(60), [&] { return w->ar_gen != g; })) { err = "emu comm: allreduce timed out"; return AZ_ERR_COMM; }
        for (int k = 0; k < n; k++) dev[k] = w->ar_out[k];
        return AZ_OK;
    }
};
// (EmuComm::share is defined below the struct: the ranks are threads of one process, a "peer mapping" is the pointer itself)
// Second transport of the emulator build: the exchange and the all-reduce are handed to callbacks, so that a
// test can carry them over torch.distributed (gloo) between real processes (tests/test_gloo_bricks.py).
typedef int (*az_emu_exchange_cb)(int n_ops, const int* peers, const void* const* send, const size_t* send_bytes,
                                  void* const* recv, const size_t* recv_bytes);
typedef int (*az_emu_allreduce_cb)(double* data, int n, int take_max);
az_emu_exchange_cb g_exchange_cb = nullptr;
az_emu_allreduce_cb g_allreduce_cb = nullptr;

struct CallbackComm : AzComm {
    int exchange(const std::vector<AzP2P>& ops, cudaStream_t) override
    {
        std::vector<int> peers; std::vector<const void*> sp; std::vector<size_t> sb, rb; std::vector<void*> rp;
        for (const AzP2P& o : ops) { peers.push_back(o.peer); sp.push_back(o.send); sb.push_back(o.send_bytes); rp.push_back(o.recv); rb.push_back(o.recv_bytes); }
        int rc = g_exchange_cb((int)ops.size(), peers.data(), sp.data(), sb.data(), rp.data(), rb.data());
        if (rc) { err = "callback comm: exchange failed"; return AZ_ERR_COMM; }
        return AZ_OK;
    }
    int allreduce(double* dev, int n, bool take_max, cudaStream_t) override
    {
        int rc = g_allreduce_cb(dev, n, take_max ? 1 : 0);
        if (rc) { err = "callback comm: allreduce failed"; return AZ_ERR_COMM; }
        return AZ_OK;
    }
};
} // namespace

extern "C" void az_emu_set_comm_callbacks(az_emu_exchange_cb ex, az_emu_allreduce_cb ar) { g_exchange_cb = ex; g_allreduce_cb = ar; }

int az_comm_make_id(void* id128, std::string*)
{
    static int counter = 0;
    std::lock_guard<std::mutex> lk(g_m);
    memset(id128, 0, 128);
    snprintf((char*)id128, 128, "emu-world-%d", ++counter);
    return AZ_OK;
}

int az_comm_create(const void* id, int rank, int nranks, int, AzComm** out, std::string* err)
{
    if (!id) { if (err) *err = "multi-rank create needs an id"; return AZ_ERR_INVALID; }
    if (g_exchange_cb && g_allreduce_cb) {
        CallbackComm* c = new CallbackComm();
        c->rank = rank; c->nranks = nranks;
        *out = c;
        return AZ_OK;
    }
    std::string key((const char*)id, strnlen((const char*)id, 128));
    std::lock_guard<std::mutex> lk(g_m);
    World*& w = g_worlds[key];
    if (!w) { w = new World(); w->nranks = nranks; }
    EmuComm* c = new EmuComm();
    c->rank = rank; c->nranks = nranks; c->w = w;
    *out = c;
    return AZ_OK;
}
