// Lennard-Jones without Coulomb (case study 1, the fcc argon benchmarks).
#include "az_pair_launch.cuh"
int az_launch_pair_lj(az_engine* e, const AzPairArgs& A, bool energy, int mode)
{
    if (e->pair_single) return az_pair_go_e<AZ_VDW_LJ, AZ_ELEC_NONE, true, false>(e, A, energy, mode);
    return az_pair_go_e<AZ_VDW_LJ, AZ_ELEC_NONE, false, false>(e, A, energy, mode);
}
