// Host build of colvars-finder_b200/csrc/cvf_math.cuh for the CPU tests.  TEST INFRASTRUCTURE ONLY.
// The header is the code the CUDA kernels run per frame (rotation from the covariance, feature stencils);
// compiling it with g++ lets tests/test_host.py check it against the numpy oracle without a GPU.
#include "../colvars-finder_b200/csrc/cvf_math.cuh"

extern "C" void host_rotation(const double* H, int n, float* R, float* Kinv) {
  for (int i = 0; i < n; ++i) cvf_rotation(H + 9 * i, R + 9 * i, Kinv + 6 * i);
}

// the same with the fp64 rotation the kernels transform with (before its rounding to float)
extern "C" void host_rotation_d(const double* H, int n, double* Rd, float* Kinv) {
  for (int i = 0; i < n; ++i) {
    float R[9];
    cvf_rotation(H + 9 * i, R, Kinv + 6 * i, Rd + 9 * i);
  }
}

extern "C" void host_dihedral(const float* p, int n, float* cs_sn, float* g) {
  for (int i = 0; i < n; ++i) {
    const float* q = p + 12 * i;
    cvf_v3 gg[4];
    float cs, sn;
    cvf_dihedral(v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]), v3(q[6], q[7], q[8]), v3(q[9], q[10], q[11]), cs, sn, gg);
    cs_sn[2 * i] = cs, cs_sn[2 * i + 1] = sn;
    for (int a = 0; a < 4; ++a) g[12 * i + 3 * a] = gg[a].x, g[12 * i + 3 * a + 1] = gg[a].y, g[12 * i + 3 * a + 2] = gg[a].z;
  }
}

extern "C" void host_angle(const float* p, int n, float* cs, float* g) {
  for (int i = 0; i < n; ++i) {
    const float* q = p + 9 * i;
    cvf_v3 ga, gc;
    cs[i] = cvf_angle(v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]), v3(q[6], q[7], q[8]), ga, gc);
    g[6 * i] = ga.x, g[6 * i + 1] = ga.y, g[6 * i + 2] = ga.z, g[6 * i + 3] = gc.x, g[6 * i + 4] = gc.y, g[6 * i + 5] = gc.z;
  }
}
