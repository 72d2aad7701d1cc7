// Instantiations of the env kernels for Crosswalk_hybrid_multi_coop (CO).
// (max car slots, max pedestrians) pairs; mhppo_env_create picks the smallest one that fits.
#include "env_kernels.cuh"
namespace mhppo {
static const EnvKernelEntry kTable[] = {
    MHPPO_ENV_ENTRY(V_COOP, 2, 1),
    MHPPO_ENV_ENTRY(V_COOP, 2, 4),
    MHPPO_ENV_ENTRY(V_COOP, 4, 4),
    MHPPO_ENV_ENTRY(V_COOP, 8, 4),
};
const EnvKernelEntry *env_table_coop(int *n) { *n = (int)(sizeof(kTable) / sizeof(kTable[0])); return kTable; }
}  // namespace mhppo
