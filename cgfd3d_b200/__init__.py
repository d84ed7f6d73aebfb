"""B200-native time-stepping hot path of CGFD3D (RK4 + curvilinear collocated-grid RHS + CFS-PML +
traction-image free surface) behind a C ABI. See DESIGN.md."""
