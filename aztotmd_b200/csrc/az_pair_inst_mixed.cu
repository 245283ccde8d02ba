// Generic path: per-pair switch over the table's kinds, any Coulomb kind, radii on demand.
#include "az_pair_launch.cuh"
int az_launch_pair_mixed(az_engine* e, const AzPairArgs& A, bool energy, int mode)
{
    if (e->pair_radii) {
        switch (e->pair_ck) {
        case AZ_ELEC_NONE: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_NONE, false, true>(e, A, energy, mode);
        case AZ_ELEC_DIRECT: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_DIRECT, false, true>(e, A, energy, mode);
        case AZ_ELEC_EWALD: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_EWALD, false, true>(e, A, energy, mode);
        default: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_FENNEL, false, true>(e, A, energy, mode);
        }
    }
    switch (e->pair_ck) {
    case AZ_ELEC_NONE: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_NONE, false, false>(e, A, energy, mode);
    case AZ_ELEC_DIRECT: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_DIRECT, false, false>(e, A, energy, mode);
    case AZ_ELEC_EWALD: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_EWALD, false, false>(e, A, energy, mode);
    default: return az_pair_go_e<AZ_VK_MIXED, AZ_ELEC_FENNEL, false, false>(e, A, energy, mode);
    }
}

int az_launch_pair(az_engine* e, const AzPairArgs& A, bool energy, int mode)
{
    switch (e->pair_vk) {
    case AZ_VDW_LJ: return az_launch_pair_lj(e, A, energy, mode);
    case AZ_VDW_BUCK: return az_launch_pair_buck(e, A, energy, mode);
    case AZ_VDW_SURK: return az_launch_pair_surk(e, A, energy, mode);
    default: return az_launch_pair_mixed(e, A, energy, mode);
    }
}
