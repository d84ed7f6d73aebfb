// one medium (CGFD_MED) x one RK stage kind (CGFD_KIND): every instantiation of the interior and free-surface kernels of that pair
#include "kernels_main.cuh"
namespace cgfd {
CGFD_INSTANTIATE_KIND(CGFD_KIND, CGFD_MED)
}
