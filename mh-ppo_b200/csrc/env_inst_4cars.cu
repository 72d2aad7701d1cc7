// Instantiations of the env kernels for Crosswalk_hybrid_multi_coop_4cars (C4); car slots = leaders + followers, exact.
// (max car slots, max pedestrians) pairs; mhppo_env_create picks the smallest one that fits.
#include "env_kernels.cuh"
namespace mhppo {
static const EnvKernelEntry kTable[] = {
    MHPPO_ENV_ENTRY(V_4CARS, 2, 2),
    MHPPO_ENV_ENTRY(V_4CARS, 4, 2),
    MHPPO_ENV_ENTRY(V_4CARS, 4, 4),
    MHPPO_ENV_ENTRY(V_4CARS, 6, 4),
    MHPPO_ENV_ENTRY(V_4CARS, 8, 4),
};
const EnvKernelEntry *env_table_4cars(int *n) { *n = (int)(sizeof(kTable) / sizeof(kTable[0])); return kTable; }
}  // namespace mhppo
