// az_pair_launch.cuh -- internal: one launch of an az_k_pair instantiation.
#pragma once
#include "az_engine.h"

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

template <int VK, int CK, bool SINGLE, bool RADII>
static int az_pair_go_e(az_engine* e, const AzPairArgs& A, bool energy)
{
    return energy ? az_pair_go<VK, CK, SINGLE, RADII, true>(e, A) : az_pair_go<VK, CK, SINGLE, RADII, false>(e, A);
}
