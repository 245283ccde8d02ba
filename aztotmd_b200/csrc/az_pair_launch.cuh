// az_pair_launch.cuh -- internal: one launch of an az_k_pair / az_k_pair_list instantiation.
#pragma once
#include "az_engine.h"

// mode: 0 = cell-tile force sweep (shared-memory windows, one thread per atom), 1 = list build + list force pass,
// 2 = list force pass, 3 = cell sweep with one warp per atom
#define AZ_PAIR_MODE_TILE 0
#define AZ_PAIR_MODE_TILE_LIST 1
#define AZ_PAIR_MODE_LIST 2
#define AZ_PAIR_MODE_WARP 3
#define AZ_PAIR_MODE_THREAD 4 // cell sweep with one thread per atom (cells much wider than the cut-off)

template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
static int az_pair_go(az_engine* e, const AzPairArgs& A)
{
#ifndef AZ_EMU
    static size_t attr_smem[16] = {};
    if (e->pair_smem > 48 * 1024 && e->pair_smem > attr_smem[e->device & 15]) {
        cudaError_t err = cudaFuncSetAttribute(az_k_pair<VK, CK, SINGLE, RADII, ENERGY>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->pair_smem);
        if (err != cudaSuccess) { e->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err); return AZ_ERR_CUDA; }
        attr_smem[e->device & 15] = e->pair_smem;
    }
#endif
    AZ_LAUNCH((az_k_pair<VK, CK, SINGLE, RADII, ENERGY>), e->pair_nblocks, AZ_PAIR_TPB, e->pair_smem, e->stream, e->P, A);
    return AZ_OK;
}

// dynamic shared memory beyond 48 KB needs the opt-in, once per kernel and device
template <class K>
static int az_smem_optin(az_engine* e, K kern, size_t smem, size_t* seen /* [16] */)
{
#ifndef AZ_EMU
    if (smem > 48 * 1024 && smem > seen[e->device & 15]) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) { e->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err); return AZ_ERR_CUDA; }
        seen[e->device & 15] = smem;
    }
#endif
    return AZ_OK;
}

// the force pass over tile lists (az_tlist.cuh): same blocks and phases as az_pair_list_go
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
static int az_pair_tlist_go(az_engine* e, const AzPairArgs& A)
{
    if (e->clear_counts_bytes && A.phase != 2 && A.phase != 3) {
        cudaError_t err = cudaMemsetAsync(A.cellCount, 0, e->clear_counts_bytes, e->stream);
        if (err != cudaSuccess) { e->err = std::string("cudaMemsetAsync: ") + cudaGetErrorString(err); return AZ_ERR_CUDA; }
    }
    static size_t seen[16] = {};
    int rc = az_smem_optin(e, az_k_pair_tlist<VK, CK, SINGLE, RADII, ENERGY>, e->tl_smem_pass, seen);
    if (rc) return rc;
    cudaStream_t st = (A.phase == 3 && e->pair_stream) ? e->pair_stream : e->stream; // fused halo: the ghost-touching blocks go to the halo stream
    if (A.phase == 3 && RADII) return AZ_OK; // (no per-atom radii on the fused-halo path)
    AZ_LAUNCH((az_k_pair_tlist<VK, CK, SINGLE, RADII, ENERGY>), e->list_nblocks, AZ_LIST_TPB, e->tl_smem_pass, st, e->P, A);
    return AZ_OK;
}

static int az_build_tlist_go(az_engine* e, const AzPairArgs& A)
{
    static size_t seen_f[16] = {}, seen_g[16] = {};
    int rc = e->list_fast ? az_smem_optin(e, az_k_build_tlist<true>, e->tl_smem_build, seen_f) : az_smem_optin(e, az_k_build_tlist<false>, e->tl_smem_build, seen_g);
    if (rc) return rc;
    if (e->list_fast) AZ_LAUNCH(az_k_build_tlist<true>, e->list_nblocks, AZ_LIST_TPB, e->tl_smem_build, e->stream, e->P, A);
    else AZ_LAUNCH(az_k_build_tlist<false>, e->list_nblocks, AZ_LIST_TPB, e->tl_smem_build, e->stream, e->P, A);
    return AZ_OK;
}

template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
static int az_pair_list_go(az_engine* e, const AzPairArgs& A)
{
    if (e->use_tlist) return az_pair_tlist_go<VK, CK, SINGLE, RADII, ENERGY>(e, A);
    // the rider of the list kernel bins the next step: clear_clist comes after everything that still reads this
    // step's counts (the list build just before) and before the kernel that counts again
    if (e->clear_counts_bytes && A.phase != 2 && A.phase != 3) {
        cudaError_t err = cudaMemsetAsync(A.cellCount, 0, e->clear_counts_bytes, e->stream);
        if (err != cudaSuccess) { e->err = std::string("cudaMemsetAsync: ") + cudaGetErrorString(err); return AZ_ERR_CUDA; }
    }
    if (A.phase == 3) { // fused halo: the ghost-touching blocks, on the halo stream (no per-atom radii on this path)
        if constexpr (!RADII) AZ_LAUNCH((az_k_pair_list<VK, CK, SINGLE, RADII, ENERGY, true>), e->list_nblocks, AZ_LIST_TPB, 0, e->pair_stream ? e->pair_stream : e->stream, e->P, A);
        return AZ_OK;
    }
    AZ_LAUNCH((az_k_pair_list<VK, CK, SINGLE, RADII, ENERGY>), e->list_nblocks, AZ_LIST_TPB, 0, e->stream, e->P, A);
    return AZ_OK;
}

template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
static int az_pair_warp_go(az_engine* e, const AzPairArgs& A)
{
    AZ_LAUNCH_PDL(e->pdl && e->profile == 0, (az_k_pair_warp<VK, CK, SINGLE, RADII, ENERGY>), e->warp_nblocks, AZ_WARP_TPB, 0, e->stream, e->P, A);
    return AZ_OK;
}

template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
static int az_pair_thread_go(az_engine* e, const AzPairArgs& A)
{
    AZ_LAUNCH((az_k_pair_thread<VK, CK, SINGLE, RADII, ENERGY>), e->thread_nblocks, AZ_THREAD_TPB, 0, e->stream, e->P, A);
    return AZ_OK;
}

template <int VK, int CK, bool SINGLE, bool RADII>
static int az_pair_go_e(az_engine* e, const AzPairArgs& A, bool energy, int mode)
{
    if (mode == AZ_PAIR_MODE_THREAD)
        return energy ? az_pair_thread_go<VK, CK, SINGLE, RADII, true>(e, A) : az_pair_thread_go<VK, CK, SINGLE, RADII, false>(e, A);
    if (mode == AZ_PAIR_MODE_WARP)
        return energy ? az_pair_warp_go<VK, CK, SINGLE, RADII, true>(e, A) : az_pair_warp_go<VK, CK, SINGLE, RADII, false>(e, A);
    if (mode == AZ_PAIR_MODE_LIST)
        return energy ? az_pair_list_go<VK, CK, SINGLE, RADII, true>(e, A) : az_pair_list_go<VK, CK, SINGLE, RADII, false>(e, A);
    if (mode == AZ_PAIR_MODE_TILE_LIST) { // record the neighbours, then take this step's forces from the list like every other step
        if (e->use_tlist) { int rc = az_build_tlist_go(e, A); if (rc) return rc; }
        else if (e->list_fast && e->list_staged) AZ_LAUNCH(az_k_build_list_staged, e->list_nblocks, AZ_LIST_TPB, 0, e->stream, e->P, A);
        else if (e->list_fast) AZ_LAUNCH(az_k_build_list<true>, e->list_nblocks, AZ_LIST_TPB, 0, e->stream, e->P, A);
        else AZ_LAUNCH(az_k_build_list<false>, e->list_nblocks, AZ_LIST_TPB, 0, e->stream, e->P, A);
        return energy ? az_pair_list_go<VK, CK, SINGLE, RADII, true>(e, A) : az_pair_list_go<VK, CK, SINGLE, RADII, false>(e, A);
    }
    return energy ? az_pair_go<VK, CK, SINGLE, RADII, true>(e, A) : az_pair_go<VK, CK, SINGLE, RADII, false>(e, A);
}
