// visco-elastic isotropic medium, generalised Maxwell body (forward/sv_curv_col_vis_iso.c)
#include "kernels_main.cuh"
namespace cgfd {
CGFD_INSTANTIATE_MEDIUM(MED_VIS)
}
