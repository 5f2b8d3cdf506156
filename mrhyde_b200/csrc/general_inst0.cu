// Part 0 of the general element kernel's instantiation list (general_dispatch.hpp: MRH_GEN_LIST_0).
#include "general_launch.cuh"

namespace mrhyde_b200 {
MRH_GEN_PART(gen_device_part0, MRH_GEN_LIST_0)
}  // namespace mrhyde_b200
