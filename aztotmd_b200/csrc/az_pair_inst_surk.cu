// Radius/temperature-dependent potential without Coulomb (case study 2).
#include "az_pair_launch.cuh"
int az_launch_pair_surk(az_engine* e, const AzPairArgs& A, bool energy, int mode)
{
    if (e->pair_single) return az_pair_go_e<AZ_VDW_SURK, AZ_ELEC_NONE, true, true>(e, A, energy, mode);
    return az_pair_go_e<AZ_VDW_SURK, AZ_ELEC_NONE, false, true>(e, A, energy, mode);
}
