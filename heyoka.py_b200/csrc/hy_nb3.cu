// Register-resident N-body kernel for 3 bodies (see hy_nbody_reg.cuh, hy_nb_launch.hpp).
#include "hy_nb_launch.hpp"

namespace hy {
HY_NB_INSTANTIATE(3)
}
