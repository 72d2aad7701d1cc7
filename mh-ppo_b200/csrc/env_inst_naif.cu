// Instantiations of the env kernels for Crosswalk_hybrid_multi_naif (NA).
// (max car slots, max pedestrians) pairs; mhppo_env_create picks the smallest one that fits.
#include "env_kernels.cuh"
namespace mhppo {
static const EnvKernelEntry kTable[] = {
    MHPPO_ENV_ENTRY(V_NAIF, 1, 2),
    MHPPO_ENV_ENTRY(V_NAIF, 2, 4),
    MHPPO_ENV_ENTRY(V_NAIF, 4, 4),
    MHPPO_ENV_ENTRY(V_NAIF, 8, 4),
};
const EnvKernelEntry *env_table_naif(int *n) { *n = (int)(sizeof(kTable) / sizeof(kTable[0])); return kTable; }
}  // namespace mhppo
