// isotropic elastic medium (forward/sv_curv_col_el_iso.c) + the scheme constants shared by every medium
#include "kernels_main.cuh"
namespace cgfd {
__constant__ FdConst c_fd;
CGFD_INSTANTIATE_MEDIUM(MED_ISO)
}
