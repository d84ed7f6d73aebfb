// VTI medium (forward/sv_curv_col_el_vti.c)
#include "kernels_main.cuh"
namespace cgfd {
CGFD_INSTANTIATE_MEDIUM(MED_VTI)
}
