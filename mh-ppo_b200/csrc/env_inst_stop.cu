// Instantiations of the env kernels for Crosswalk_hybrid_multi_stop (ST).
// (max car slots, max pedestrians) pairs; mhppo_env_create picks the smallest one that fits.
#include "env_kernels.cuh"
namespace mhppo {
static const EnvKernelEntry kTable[] = {
    MHPPO_ENV_ENTRY(V_STOP, 1, 2),
    MHPPO_ENV_ENTRY(V_STOP, 2, 4),
    MHPPO_ENV_ENTRY(V_STOP, 4, 4),
    MHPPO_ENV_ENTRY(V_STOP, 8, 4),
};
const EnvKernelEntry *env_table_stop(int *n) { *n = (int)(sizeof(kTable) / sizeof(kTable[0])); return kTable; }
}  // namespace mhppo
