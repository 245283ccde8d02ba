// Buckingham with every real-space Coulomb kind (the ionic benchmark).
#include "az_pair_launch.cuh"
int az_launch_pair_buck(az_engine* e, const AzPairArgs& A, bool energy, int mode)
{
    switch (e->pair_ck) {
    case AZ_ELEC_NONE: return az_pair_go_e<AZ_VDW_BUCK, AZ_ELEC_NONE, false, false>(e, A, energy, mode);
    case AZ_ELEC_DIRECT: return az_pair_go_e<AZ_VDW_BUCK, AZ_ELEC_DIRECT, false, false>(e, A, energy, mode);
    case AZ_ELEC_EWALD: return az_pair_go_e<AZ_VDW_BUCK, AZ_ELEC_EWALD, false, false>(e, A, energy, mode);
    default: return az_pair_go_e<AZ_VDW_BUCK, AZ_ELEC_FENNEL, false, false>(e, A, energy, mode);
    }
}
