// general anisotropic medium (forward/sv_curv_col_el_aniso.c)
#include "kernels_main.cuh"
namespace cgfd {
CGFD_INSTANTIATE_MEDIUM(MED_ANISO)
}
