// Part 1 of the general element kernel's instantiation list (general_dispatch.hpp: MRH_GEN_LIST_1).
#include "general_launch.cuh"

namespace mrhyde_b200 {
MRH_GEN_PART(gen_device_part1, MRH_GEN_LIST_1)
}  // namespace mrhyde_b200
