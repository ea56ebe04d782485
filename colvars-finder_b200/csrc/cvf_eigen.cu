// cvf_eigen.cu -- EigenFunctionTask.loss_func (generator branch) + backward on sm_100a.
//
// Replaces, per mini-batch (reference file:line under /root/reference/colvarsfinder):
//   core.py:403      y = model(pp_layer(X))                          -> load_tile / preprocess_frame / forward phases
//   core.py:406-410  total weight, weighted means / variances         -> fp64 batch sums S0, S1, S2
//   core.py:424      torch.autograd.grad(y_i.sum(), X, create_graph)  -> reverse phases + closed-form J_r^T (jphase)
//   core.py:426,438  Rayleigh quotients / objective                   -> SD + cvf_eigen_combine
//   core.py:446-455  penalty, total loss                              -> cvf_eigen_combine
//   core.py:517      loss.backward() (double backward)                -> pass 2: one tangent sweep + one reverse sweep
//
// One persistent CTA per SM walks over tiles of F frames.  See cvf_common.cuh for the row layout.  The k
// networks are processed one after the other inside a tile so that one network's state
// (activations A, adjoints G, tangents T, second adjoints S) fits in shared memory next to the frame.
#include <atomic>
#include <math.h>
#include <string.h>

#include "cvf_common.cuh"
#include "cvf_math.cuh"

namespace cvf {

// cvf_eigen_fast.cu: thread-private / FFMA2 path for the common network shape
bool fast_eigen_supported(const cvf_preproc* pp, const NetPlan& np, int k);
size_t fast_eigen_workspace_bytes(const cvf_preproc* pp, const NetPlan& np, int k, long long B);
int fast_eigen_stats(const cvf_preproc* pp, const NetPlan& np, int k, const float* x, const float* w, long long B,
                     const float* params, float* y_out, double* stats_out, void* workspace, size_t ws_bytes, cudaStream_t stream);
int fast_eigen_grad(const cvf_preproc* pp, const NetPlan& np, int k, const float* x, const float* w, long long B,
                    const float* params, const double* combine, const float* seed_extra, double* grad_out, void* workspace,
                    size_t ws_bytes, int scratch_valid, cudaStream_t stream);
static std::atomic<int> g_eigen_path{0};   // 0: fast path whenever it applies, 1: always the general row-engine kernels

// Optional per-phase cycle counters (profiling builds only: -DCVF_PHASE_TIMERS, see profiles/phase_timing.py).
#ifdef CVF_PHASE_TIMERS
__device__ unsigned long long g_phase_cycles[16];
#define PT_DECL long long pt_last = clock64(); unsigned long long pt_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define PT_MARK(slot)                          \
  do {                                         \
    if (threadIdx.x == 0) {                    \
      const long long t_ = clock64();          \
      pt_acc[slot] += t_ - pt_last;            \
      pt_last = t_;                            \
    }                                          \
  } while (0)
#define PT_FLUSH                                                                        \
  do {                                                                                  \
    if (threadIdx.x == 0)                                                               \
      for (int i_ = 0; i_ < 12; ++i_) atomicAdd(&g_phase_cycles[i_], pt_acc[i_]);       \
  } while (0)
#else
#define PT_DECL
#define PT_MARK(slot)
#define PT_FLUSH
#endif
// slots: 0 load, 1 preprocess, 2 forward, 3 reverse, 4 J phase, 5 tangent, 6 outer(tangent part), 7 second reverse,
//        8 outer(all layers), 9 stats, 10 setup

struct EigenPlan {
  NetPlan net;
  int k;
  int F, FS, FB;  // frames per tile, row stride, 4-frame blocks per tile
  int nthreads;
  // pre-processing
  int kind, dim, n_atoms, n_used, n_align, n_feat, d_r;
  const int32_t* used_atoms;
  const int32_t* align_used;
  const float* ref;
  const int32_t* feat;
  const float* diag;
  int pos_alias;   // features are exactly the positions of the used atoms in order: r rows alias the y rows
  int used_identity;   // used_atoms = 0..n_atoms-1: a tile is one contiguous block of global memory
  // rows (units of FS floats from the row base)
  int row_w, row_one, row_seed, row_R, row_Kinv, row_y, row_D, row_Y, row_r, row_net;
  int a_row[kMaxLayers + 1], g_row[kMaxLayers + 1], t_row[kMaxLayers + 1], s_row[kMaxLayers + 1];  // relative to row_net
  int v_row, gx_row;   // relative to row_net
  int n_rows;
  // shared memory carve-up (floats)
  int off_params, off_rows, off_comb, off_red;
  size_t smem_bytes;
  int n_stats;
};

static int build_plan(const cvf_preproc* pp, const cvf_mlp* net, int k, EigenPlan* P) {
  memset(P, 0, sizeof(*P));
  if (!pp || !net) {
    set_error("null descriptor");
    return CVF_E_ARG;
  }
  if (k < 1 || k > kMaxK) {
    set_error("k = %d outside [1,%d]", k, kMaxK);
    return CVF_E_UNSUPPORTED;
  }
  int e = make_net_plan(net, &P->net);
  if (e) return e;
  const NetPlan& np = P->net;
  if (np.dims[np.L] != 1) {
    set_error("eigenfunction networks must be scalar valued (nn.py:270)");
    return CVF_E_ARG;
  }
  for (int l = 0; l < np.L; ++l)
    if ((l < np.L - 1) ? (np.act[l] == CVF_ACT_NONE || np.act[l] != np.act[0]) : (np.act[l] != CVF_ACT_NONE)) {
      set_error("eigenfunction networks: one activation after every layer but the last");
      return CVF_E_UNSUPPORTED;
    }
  P->k = k;
  P->kind = pp->kind;
  if (pp->kind == 0) {
    if (pp->dim < 1 || pp->dim != np.dims[0]) {
      set_error("identity pre-processing: dim %d != network input %d", pp->dim, np.dims[0]);
      return CVF_E_ARG;
    }
    P->dim = pp->dim;
    P->d_r = pp->dim;
  } else if (pp->kind == 1) {
    if (pp->n_atoms < 1 || pp->n_used < 1 || pp->n_used > pp->n_atoms || !pp->used_atoms || pp->n_feat < 1 || !pp->feat ||
        pp->d_r != np.dims[0] || (pp->n_align > 0 && (!pp->align_used || !pp->ref)) || pp->n_align < 0) {
      set_error("molecular pre-processing descriptor inconsistent (d_r %d, network input %d)", pp->d_r, np.dims[0]);
      return CVF_E_ARG;
    }
    if (pp->n_align > 0 && pp->n_align < 3) {
      set_error("alignment needs at least 3 atoms");
      return CVF_E_ARG;
    }
    P->n_atoms = pp->n_atoms, P->n_used = pp->n_used, P->n_align = pp->n_align, P->n_feat = pp->n_feat, P->d_r = pp->d_r;
    P->used_atoms = pp->used_atoms, P->align_used = pp->align_used, P->ref = pp->ref, P->feat = pp->feat;
    P->used_identity = pp->used_identity ? 1 : 0;
    if (pp->positions_only && (pp->n_feat != pp->n_used || pp->d_r != 3 * pp->n_used)) {
      set_error("positions_only needs n_feat == n_used and d_r == 3 n_used");
      return CVF_E_ARG;
    }
  } else {
    set_error("unknown pre-processing kind %d", pp->kind);
    return CVF_E_ARG;
  }
  P->diag = pp->diag;
  P->n_stats = 1 + 2 * k + k * k;
  return 0;
}

// Decide the row map for a given tile size; returns the shared-memory bytes needed.
static size_t layout_plan(EigenPlan* P, int F, bool pos_alias) {
  const NetPlan& np = P->net;
  P->F = F, P->FS = F + 4, P->FB = F / 4;
  P->pos_alias = pos_alias ? 1 : 0;
  int r = 0;
  P->row_w = r++;
  P->row_one = r++;
  P->row_seed = r++;
  P->row_R = r, r += 9;
  P->row_Kinv = r, r += 6;
  P->row_y = r, r += P->k;
  P->row_D = r, r += P->k;
  if (P->kind == 1) {
    P->row_Y = r, r += 3 * P->n_used;
    if (pos_alias) P->row_r = P->row_Y;
    else P->row_r = r, r += P->d_r;
  } else {
    P->row_Y = r;
    P->row_r = r, r += P->d_r;
  }
  P->row_net = r;
  int sumH = 0;
  for (int l = 1; l < np.L; ++l) sumH += np.dims[l];
  int q = 0;
  for (int l = 1; l < np.L; ++l) P->a_row[l] = q, q += np.dims[l];
  for (int l = 1; l < np.L; ++l) P->g_row[l] = q, q += np.dims[l];
  // region shared by (T,S) and, before the tangent sweep starts, by V (d_r rows) and GX (3 n_used rows):
  //   [ T_1 | S_1 | T_2 .. | S_2 .. ]   and   [ T_1 | S_1 | V | GX ]
  const int ts0 = q;
  int t = ts0;
  if (np.L > 1) {
    P->t_row[1] = t, t += np.dims[1];
    P->s_row[1] = t, t += np.dims[1];
  }
  const int after1 = t;
  for (int l = 2; l < np.L; ++l) P->t_row[l] = t, t += np.dims[l];
  for (int l = 2; l < np.L; ++l) P->s_row[l] = t, t += np.dims[l];
  P->v_row = after1;
  int vend = after1 + P->d_r;
  if (P->kind == 1 && !pos_alias) {
    P->gx_row = vend;
    vend += 3 * P->n_used;
  } else {
    P->gx_row = P->v_row;   // positions only: dg/dy is u itself and dy is v itself
  }
  q = t > vend ? t : vend;
  (void)sumH;
  P->n_rows = P->row_net + q;
  // floats
  int off = 0;
  P->off_params = off, off += P->k * np.smem_floats;
  P->off_comb = off, off += round4(2 * (3 + 5 * P->k + P->k * P->k));    // combine vector as doubles
  P->off_red = off, off += 2 * 4 * P->n_stats;   // warp-reduction scratch: n_stats x 4 warps, doubles
  off = (off + 3) & ~3;
  P->off_rows = off, off += P->n_rows * P->FS;
  P->smem_bytes = (size_t)off * sizeof(float);
  return P->smem_bytes;
}

static int finish_plan(EigenPlan* P, bool pos_alias) {
  const size_t cap = (size_t)max_smem_optin();
  static const int Fs[3] = {128, 64, 32};
  for (int c = 0; c < 3; ++c) {
    if (layout_plan(P, Fs[c], pos_alias) <= cap) break;
    if (c == 2) {
      set_error("network/pre-processing state does not fit shared memory (%zu B needed at 32 frames, %zu available)",
                P->smem_bytes, cap);
      return CVF_E_UNSUPPORTED;
    }
  }
  P->nthreads = 384;
  return 0;
}

// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ cvf_v3 ldv(const float* rows, int FS, int atom, int f) {
  return v3(rows[(3 * atom + 0) * FS + f], rows[(3 * atom + 1) * FS + f], rows[(3 * atom + 2) * FS + f]);
}
__device__ __forceinline__ void stv(float* rows, int FS, int atom, int f, cvf_v3 v) {
  rows[(3 * atom + 0) * FS + f] = v.x;
  rows[(3 * atom + 1) * FS + f] = v.y;
  rows[(3 * atom + 2) * FS + f] = v.z;
}
__device__ __forceinline__ void addv(float* rows, int FS, int atom, int f, cvf_v3 v) {
  rows[(3 * atom + 0) * FS + f] += v.x;
  rows[(3 * atom + 1) * FS + f] += v.y;
  rows[(3 * atom + 2) * FS + f] += v.z;
}

// Kabsch alignment of frame f in place on the Y rows + feature values into the r rows (thread per frame).
__device__ void preprocess_frame(const EigenPlan& P, float* rows, int f) {
  const int FS = P.FS;
  float* Y = rows + P.row_Y * FS;
  if (P.n_align > 0) {
    double cx = 0, cy = 0, cz = 0;
    for (int a = 0; a < P.n_align; ++a) {
      const cvf_v3 p = ldv(Y, FS, P.align_used[a], f);
      cx += p.x, cy += p.y, cz += p.z;
    }
    const double inv = 1.0 / P.n_align;
    cx *= inv, cy *= inv, cz *= inv;
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < P.n_align; ++a) {
      const cvf_v3 p = ldv(Y, FS, P.align_used[a], f);
      const double px = p.x - cx, py = p.y - cy, pz = p.z - cz;
      const double rx = P.ref[3 * a], ry = P.ref[3 * a + 1], rz = P.ref[3 * a + 2];
      H[0] += px * rx, H[1] += px * ry, H[2] += px * rz;
      H[3] += py * rx, H[4] += py * ry, H[5] += py * rz;
      H[6] += pz * rx, H[7] += pz * ry, H[8] += pz * rz;
    }
    float R[9], Ki[6];
    double Rd[9];
    cvf_rotation(H, R, Ki, Rd);
#pragma unroll
    for (int i = 0; i < 9; ++i) rows[(P.row_R + i) * FS + f] = R[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) rows[(P.row_Kinv + i) * FS + f] = Ki[i];
    for (int a = 0; a < P.n_used; ++a) {
      const cvf_v3 p = ldv(Y, FS, a, f);
      stv(Y, FS, a, f, cvf_transform(p.x, p.y, p.z, cx, cy, cz, Rd));
    }
  }
  if (P.pos_alias) return;
  float* r = rows + P.row_r * FS;
  int col = 0;
  for (int j = 0; j < P.n_feat; ++j) {
    const int32_t* fr = P.feat + 5 * j;
    const int type = fr[0];
    if (type == CVF_FEAT_POSITION) {
      const cvf_v3 p = ldv(Y, FS, fr[1], f);
      r[(col + 0) * FS + f] = p.x, r[(col + 1) * FS + f] = p.y, r[(col + 2) * FS + f] = p.z;
      col += 3;
    } else if (type == CVF_FEAT_BOND) {
      cvf_v3 g;
      r[col * FS + f] = cvf_bond(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), g);
      col += 1;
    } else if (type == CVF_FEAT_ANGLE) {
      cvf_v3 ga, gc;
      r[col * FS + f] = cvf_angle(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), ldv(Y, FS, fr[3], f), ga, gc);
      col += 1;
    } else {
      float cs, sn;
      cvf_v3 g[4];
      cvf_dihedral(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), ldv(Y, FS, fr[3], f), ldv(Y, FS, fr[4], f), cs, sn, g);
      r[col * FS + f] = cs, r[(col + 1) * FS + f] = sn;
      col += 2;
    }
  }
}

// J phase for frame f of one network.  In: V rows = u = dg/dr.  Out: returns the Dirichlet density
// D = sum_j a_j (df/dx_j)^2;  if `tangent`, V rows <- scale * J_r (a .* J_r^T u).
__device__ float jphase_frame(const EigenPlan& P, float* rows, float* netrows, int f, bool tangent, float scale) {
  const int FS = P.FS;
  float* V = netrows + P.v_row * FS;
  float D = 0.0f;
  if (P.kind == 0) {
    for (int j = 0; j < P.dim; ++j) {
      const float u = V[j * FS + f];
      const float a = P.diag ? P.diag[j] : 1.0f;
      D = fmaf(a * u, u, D);
      if (tangent) V[j * FS + f] = scale * a * u;
    }
    return D;
  }
  const float* Y = rows + P.row_Y * FS;
  float* GX = netrows + P.gx_row * FS;
  // 1. G = (d r / d y)^T u
  if (!P.pos_alias) {
    for (int j = 0; j < 3 * P.n_used; ++j) GX[j * FS + f] = 0.0f;
    int col = 0;
    for (int j = 0; j < P.n_feat; ++j) {
      const int32_t* fr = P.feat + 5 * j;
      const int type = fr[0];
      if (type == CVF_FEAT_POSITION) {
        addv(GX, FS, fr[1], f, v3(V[col * FS + f], V[(col + 1) * FS + f], V[(col + 2) * FS + f]));
        col += 3;
      } else if (type == CVF_FEAT_BOND) {
        cvf_v3 g;
        cvf_bond(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), g);
        const float u = V[col * FS + f];
        addv(GX, FS, fr[2], f, u * g);
        addv(GX, FS, fr[1], f, (-u) * g);
        col += 1;
      } else if (type == CVF_FEAT_ANGLE) {
        cvf_v3 ga, gc;
        cvf_angle(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), ldv(Y, FS, fr[3], f), ga, gc);
        const float u = V[col * FS + f];
        addv(GX, FS, fr[1], f, u * ga);
        addv(GX, FS, fr[3], f, u * gc);
        addv(GX, FS, fr[2], f, (-u) * (ga + gc));
        col += 1;
      } else {
        float cs, sn;
        cvf_v3 g[4];
        cvf_dihedral(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), ldv(Y, FS, fr[3], f), ldv(Y, FS, fr[4], f), cs, sn, g);
        const float s = -sn * V[col * FS + f] + cs * V[(col + 1) * FS + f];
#pragma unroll
        for (int a = 0; a < 4; ++a) addv(GX, FS, fr[1 + a], f, s * g[a]);
        col += 2;
      }
    }
  }
  // 2. through the alignment, weight by diag_coeff, and back (J_r of the alignment)
  float R[9], Ki[6];
  if (P.n_align > 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = rows[(P.row_R + i) * FS + f];
#pragma unroll
    for (int i = 0; i < 6; ++i) Ki[i] = rows[(P.row_Kinv + i) * FS + f];
    cvf_v3 tau = v3(0, 0, 0), gs = v3(0, 0, 0);
    for (int a = 0; a < P.n_used; ++a) {
      const cvf_v3 g = ldv(GX, FS, a, f);
      tau = tau + cross(g, ldv(Y, FS, a, f));
      gs = gs + g;
    }
    const cvf_v3 q = mul_sym(Ki, tau);
    const cvf_v3 gm = (1.0f / P.n_align) * gs;
    for (int a = 0; a < P.n_align; ++a) {
      const int at = P.align_used[a];
      const cvf_v3 rf = v3(P.ref[3 * a], P.ref[3 * a + 1], P.ref[3 * a + 2]);
      addv(GX, FS, at, f, (-1.0f) * (gm + cross(rf, q)));
    }
  }
  // GX now holds  (df/dx) R  (i.e. the x-space gradient rotated into the aligned frame)
  if (P.diag == nullptr) {
    for (int a = 0; a < P.n_used; ++a) {
      const cvf_v3 g = ldv(GX, FS, a, f);
      D += dot(g, g);
    }
  } else {
    for (int a = 0; a < P.n_used; ++a) {
      cvf_v3 g = ldv(GX, FS, a, f);
      if (P.n_align > 0) g = mul_rowvec_T(g, R);          // df/dx
      const cvf_v3 h = v3(P.diag[3 * a] * g.x, P.diag[3 * a + 1] * g.y, P.diag[3 * a + 2] * g.z);
      D += dot(g, h);
      if (tangent) stv(GX, FS, a, f, P.n_align > 0 ? mul_rowvec(h, R) : h);   // h R
    }
  }
  if (!tangent) return D;
  if (P.n_align > 0) {
    cvf_v3 dc = v3(0, 0, 0);
    for (int a = 0; a < P.n_align; ++a) dc = dc + ldv(GX, FS, P.align_used[a], f);
    dc = (1.0f / P.n_align) * dc;
    cvf_v3 tq = v3(0, 0, 0);
    for (int a = 0; a < P.n_align; ++a) {
      const cvf_v3 rf = v3(P.ref[3 * a], P.ref[3 * a + 1], P.ref[3 * a + 2]);
      tq = tq + cross(ldv(GX, FS, P.align_used[a], f) - dc, rf);
    }
    const cvf_v3 om = mul_sym(Ki, tq);
    for (int a = 0; a < P.n_used; ++a) {
      const cvf_v3 u = ldv(GX, FS, a, f) - dc;
      stv(GX, FS, a, f, scale * (u + cross(om, ldv(Y, FS, a, f))));
    }
  } else {
    for (int j = 0; j < 3 * P.n_used; ++j) GX[j * FS + f] *= scale;
  }
  // 3. v = (d r / d y) dy
  if (!P.pos_alias) {
    int col = 0;
    for (int j = 0; j < P.n_feat; ++j) {
      const int32_t* fr = P.feat + 5 * j;
      const int type = fr[0];
      if (type == CVF_FEAT_POSITION) {
        const cvf_v3 d = ldv(GX, FS, fr[1], f);
        V[col * FS + f] = d.x, V[(col + 1) * FS + f] = d.y, V[(col + 2) * FS + f] = d.z;
        col += 3;
      } else if (type == CVF_FEAT_BOND) {
        cvf_v3 g;
        cvf_bond(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), g);
        V[col * FS + f] = dot(g, ldv(GX, FS, fr[2], f) - ldv(GX, FS, fr[1], f));
        col += 1;
      } else if (type == CVF_FEAT_ANGLE) {
        cvf_v3 ga, gc;
        cvf_angle(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), ldv(Y, FS, fr[3], f), ga, gc);
        const cvf_v3 dm = ldv(GX, FS, fr[2], f);
        V[col * FS + f] = dot(ga, ldv(GX, FS, fr[1], f) - dm) + dot(gc, ldv(GX, FS, fr[3], f) - dm);
        col += 1;
      } else {
        float cs, sn;
        cvf_v3 g[4];
        cvf_dihedral(ldv(Y, FS, fr[1], f), ldv(Y, FS, fr[2], f), ldv(Y, FS, fr[3], f), ldv(Y, FS, fr[4], f), cs, sn, g);
        float dphi = 0.0f;
#pragma unroll
        for (int a = 0; a < 4; ++a) dphi += dot(g[a], ldv(GX, FS, fr[1 + a], f));
        V[col * FS + f] = -sn * dphi, V[(col + 1) * FS + f] = cs * dphi;
        col += 2;
      }
    }
  }
  return D;
}

// ------------------------------------------------------------------------------------------------------
// Bring one tile of frames into the [unit][frame] rows.  All loads of a thread are issued before the first store
// (8 in flight per thread) and the NEXT tile of this CTA is prefetched into L2, so the HBM latency is paid once per
// tile instead of once per element.
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ void load_tile(const EigenPlan& P, const float* __restrict__ x, long long B, long long f_base, float* rows,
                          int tid, int nt) {
  const int FS = P.FS, F = P.F;
  const int fl = P.kind == 0 ? P.dim : 3 * P.n_atoms;        // floats per frame in global memory
  const int row0 = P.kind == 0 ? P.row_r : P.row_Y;
  const bool full = f_base + F <= B;
  const float* src = x + (size_t)f_base * fl;
  if ((P.kind == 0 || P.used_identity) && full && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((F * fl) & 3) == 0) {
    // contiguous tile: coalesced 16-byte loads, scattered into the transposed rows
    const int nvec = (F * fl) >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (int base = 0; base < nvec; base += 8 * nt) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = base + j * nt + tid;
        if (idx < nvec) v[j] = __ldg(src4 + idx);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = base + j * nt + tid;
        if (idx < nvec) {
          const int e = 4 * idx;
          int f = e / fl, u = e - f * fl;
          const float vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            rows[(row0 + u) * FS + f] = vv[c];
            if (++u == fl) u = 0, ++f;
          }
        }
      }
    }
  } else {
    const int nu = P.kind == 0 ? P.dim : 3 * P.n_used;
    const int total = F * nu;
    for (int base = 0; base < total; base += 8 * nt) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = base + j * nt + tid;
        if (idx < total) {
          const int f = idx / nu, u = idx - f * nu;
          const long long fr = min(f_base + f, B - 1);
          int col = u;
          if (P.kind == 1) {
            const int a = u / 3;
            col = 3 * P.used_atoms[a] + (u - 3 * a);
          }
          v[j] = __ldg(x + (size_t)fr * fl + col);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = base + j * nt + tid;
        if (idx < total) {
          const int f = idx / nu, u = idx - f * nu;
          rows[(row0 + u) * FS + f] = v[j];
        }
      }
    }
  }
}

template <bool GRAD, int FPL>
__global__ void __launch_bounds__(384, 1)
eigen_kernel(const EigenPlan P, const float* __restrict__ x, const float* __restrict__ w, long long B,
             const float* __restrict__ params, float* __restrict__ y_io, const double* __restrict__ combine,
             double* __restrict__ partial, const float* __restrict__ seed_extra) {
  extern __shared__ __align__(16) float smem[];
  const NetPlan& np = P.net;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int FS = P.FS, F = P.F, k = P.k;
  const int FB = F / FPL;      // lane-sized frame blocks per tile
  typedef FVec<FPL> V;
  float* Wsm = smem + P.off_params;
  float* rows = smem + P.off_rows;
  double* comb = reinterpret_cast<double*>(smem + P.off_comb);
  double* red = reinterpret_cast<double*>(smem + P.off_red);

  for (int i = 0; i < k; ++i) load_net_params(np, params + (size_t)i * np.n_params, Wsm + i * np.smem_floats, tid, nt);
  for (int f = tid; f < FS; f += nt) rows[P.row_one * FS + f] = 1.0f;
  const int n_part = GRAD ? k * np.n_params : P.n_stats;
  double* part = partial + (size_t)blockIdx.x * n_part;
  if (GRAD) {
    const int nc = 3 + 5 * k + k * k;
    for (int i = tid; i < nc; i += nt) comb[i] = combine[i];
    for (int i = tid; i < n_part; i += nt) part[i] = 0.0;
  }
  double stat_acc = 0.0;   // thread `tid` owns batch sum number `tid` (stats pass)
  PT_DECL;
  __syncthreads();
  PT_MARK(10);
  // combine vector layout: loss, obj, pen, eig[k], cvec[k], mean[k], cD[k], C2[k*k], a0[k]
  const double* c_mean = comb + 3 + 2 * k;
  const double* c_cD = comb + 3 + 3 * k;
  const double* c_C2 = comb + 3 + 4 * k;
  const double* c_a0 = comb + 3 + 4 * k + k * k;
  const int fl = P.kind == 0 ? P.dim : 3 * P.n_atoms;

  const long long n_tiles = (B + F - 1) / F;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long f_base = tile * F;
    // ---- load the tile: frames -> rows (transposed), weights, (pass 2) y of every network
    load_tile(P, x, B, f_base, rows, tid, nt);
    for (int f = tid; f < F; f += nt) {
      const long long fr = f_base + f;
      rows[P.row_w * FS + f] = fr < B ? w[fr] : 0.0f;
      if (GRAD)
        for (int i = 0; i < k; ++i) rows[(P.row_y + i) * FS + f] = y_io[(size_t)i * B + min(fr, B - 1)];
    }
    if (tid == 0) {
      const long long nxt = tile + gridDim.x;
      if (nxt < n_tiles) {
        const long long nb = nxt * F;
        const long long nf = min((long long)F, B - nb);
        const float* pn = x + (size_t)nb * fl;
        const unsigned bytes = (unsigned)((nf * fl * 4) & ~15LL);
        if (((reinterpret_cast<uintptr_t>(pn) & 15) == 0) && bytes > 0) prefetch_l2(pn, bytes);
      }
    }
    __syncthreads();
    PT_MARK(0);
    if (P.kind == 1) {
      if (tid < F) preprocess_frame(P, rows, tid);
      __syncthreads();
      PT_MARK(1);
    }
    float* netrows = rows + P.row_net * FS;
    const float* r_in = rows + P.row_r * FS;

    const int akind = np.L > 1 ? np.act[0] : CVF_ACT_NONE;   // the one activation of the chain (nn.py:29,56)
    for (int n = 0; n < k; ++n) {
      const float* Wn = Wsm + n * np.smem_floats;
      // ---- forward (nn.py:52-57): A_l = tanh(W_l A_{l-1} + b_l), y = W_L A_{L-1} + b_L
      for (int l = 0; l < (GRAD ? np.L - 1 : np.L); ++l) {   // pass 2 keeps the y of pass 1
        const int nin = np.dims[l], nout = np.dims[l + 1];
        const float* in = l == 0 ? r_in : netrows + P.a_row[l] * FS;
        const bool last = l == np.L - 1;
        float* out = last ? rows + (P.row_y + n) * FS : netrows + P.a_row[l + 1] * FS;
        const int items = ((nout + 3) >> 2) * FB;
        for (int it = tid; it < items; it += nt) {
          const int ob = it / FB, fb = it - ob * FB;
          float acc[4][FPL];
          tile_fwd<FPL>(acc, Wn + np.w_off[l], np.ld[l], 4 * ob, nout, nin, in, FS, FPL * fb);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int o = 4 * ob + j;
            if (o < nout) {
              const float b = Wn[np.b_off[l] + o];
              V v;
#pragma unroll
              for (int f = 0; f < FPL; ++f) v.v[f] = last ? acc[j][f] + b : cvf_act(akind, acc[j][f] + b);
              v.st(out + o * FS + FPL * fb);
            }
          }
        }
        __syncthreads();
      }
      PT_MARK(2);
      // ---- reverse: G_l = adjoint of z_l for the seed dy = 1;  u = dy/dr -> V rows
      for (int l = np.L - 1; l >= 0; --l) {
        const int nin = np.dims[l], nout = np.dims[l + 1];
        const float* in = l == np.L - 1 ? rows + P.row_one * FS : netrows + P.g_row[l + 1] * FS;
        float* out = l == 0 ? netrows + P.v_row * FS : netrows + P.g_row[l] * FS;
        const float* A = l == 0 ? nullptr : netrows + P.a_row[l] * FS;
        const int items = ((nin + 3) >> 2) * FB;
        for (int it = tid; it < items; it += nt) {
          const int ib = it / FB, fb = it - ib * FB;
          float acc[4][FPL];
          tile_tr<FPL>(acc, Wn + np.w_off[l], np.ld[l], 4 * ib, nout, in, FS, FPL * fb);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * ib + j;
            if (i < nin) {
              V v;
              if (l > 0) {
                const V a = V::ld(A + i * FS + FPL * fb);
#pragma unroll
                for (int f = 0; f < FPL; ++f) v.v[f] = acc[j][f] * cvf_act_d1(akind, a.v[f]);
              } else {
#pragma unroll
                for (int f = 0; f < FPL; ++f) v.v[f] = acc[j][f];
              }
              v.st(out + i * FS + FPL * fb);
            }
          }
        }
        __syncthreads();
      }
      PT_MARK(3);
      // ---- J phase (thread per frame): Dirichlet density; pass 2: tangent direction and output seed
      if (tid < F) {
        const int f = tid;
        float scale = 0.0f;
        if (GRAD) {
          const float wf = rows[P.row_w * FS + f];
          scale = (float)(2.0 * (double)wf * c_cD[n]);
          double s = c_a0[n];
          for (int j = 0; j < k; ++j) s += c_C2[n * k + j] * ((double)rows[(P.row_y + j) * FS + f] - c_mean[j]);
          float sd = (float)((double)wf * s);
          if (seed_extra != nullptr && f_base + f < B) sd += seed_extra[(size_t)n * B + f_base + f];   // transfer-operator term
          rows[P.row_seed * FS + f] = sd;
        }
        const float D = jphase_frame(P, rows, netrows, f, GRAD, scale);
        if (!GRAD) rows[(P.row_D + n) * FS + f] = D;
      }
      __syncthreads();
      PT_MARK(4);
      if (GRAD) {
        double* pn = part + (size_t)n * np.n_params;
        // ---- tangent sweep along v:  T_l = (1-A_l^2) zdot_l,  S_l = -2 A_l G_l zdot_l   (zdot_l = W_l T_{l-1})
        for (int l = 0; l < np.L - 1; ++l) {
          const int nin = np.dims[l], nout = np.dims[l + 1];
          const float* in = l == 0 ? netrows + P.v_row * FS : netrows + P.t_row[l] * FS;
          float* T = netrows + P.t_row[l + 1] * FS;
          float* S = netrows + P.s_row[l + 1] * FS;
          const float* A = netrows + P.a_row[l + 1] * FS;
          const float* G = netrows + P.g_row[l + 1] * FS;
          const int items = ((nout + 3) >> 2) * FB;
          for (int it = tid; it < items; it += nt) {
            const int ob = it / FB, fb = it - ob * FB;
            float acc[4][FPL];
            tile_fwd<FPL>(acc, Wn + np.w_off[l], np.ld[l], 4 * ob, nout, nin, in, FS, FPL * fb);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = 4 * ob + j;
              if (o < nout) {
                const V a = V::ld(A + o * FS + FPL * fb), g = V::ld(G + o * FS + FPL * fb);
                V t, e;
#pragma unroll
                for (int f = 0; f < FPL; ++f) {
                  t.v[f] = cvf_act_d1(akind, a.v[f]) * acc[j][f];
                  e.v[f] = cvf_act_r2(akind, a.v[f]) * g.v[f] * acc[j][f];
                }
                t.st(T + o * FS + FPL * fb);
                e.st(S + o * FS + FPL * fb);
              }
            }
          }
          if (l == 0) {
            // tangent part of dW_1 while V is still alive: dW_1 += G_1 (x) v   (same inputs as the items above: no barrier)
            outer_layer(0, tid, nt, np.dims[1], np.dims[0], netrows + P.g_row[1] * FS, netrows + P.v_row * FS, nullptr, nullptr,
                        FS, F, pn + np.gw_off[0], nullptr);
          }
          __syncthreads();
        }
        if (np.L == 1) {
          outer_layer(0, tid, nt, np.dims[1], np.dims[0], rows + P.row_one * FS, netrows + P.v_row * FS, nullptr, nullptr, FS, F,
                      pn + np.gw_off[0], nullptr);
          __syncthreads();
        }
        PT_MARK(5);
        // ---- second reverse sweep: s_l (adjoint of z_l) in place of S_l;  s_L = seed
        for (int l = np.L - 1; l >= 1; --l) {
          const int nin = np.dims[l], nout = np.dims[l + 1];
          const float* in = l == np.L - 1 ? rows + P.row_seed * FS : netrows + P.s_row[l + 1] * FS;
          float* S = netrows + P.s_row[l] * FS;
          const float* A = netrows + P.a_row[l] * FS;
          const int items = ((nin + 3) >> 2) * FB;
          for (int it = tid; it < items; it += nt) {
            const int ib = it / FB, fb = it - ib * FB;
            float acc[4][FPL];
            tile_tr<FPL>(acc, Wn + np.w_off[l], np.ld[l], 4 * ib, nout, in, FS, FPL * fb);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = 4 * ib + j;
              if (i < nin) {
                const V a = V::ld(A + i * FS + FPL * fb);
                V e = V::ld(S + i * FS + FPL * fb);
#pragma unroll
                for (int f = 0; f < FPL; ++f) e.v[f] = fmaf(acc[j][f], cvf_act_d1(akind, a.v[f]), e.v[f]);
                e.st(S + i * FS + FPL * fb);
              }
            }
          }
          __syncthreads();
        }
        PT_MARK(7);
        // ---- parameter gradients: dW_l += s_l (x) A_{l-1} + G_l (x) T_{l-1},  db_l += sum_f s_l
        {
          int rot = 0;
          for (int l = 0; l < np.L; ++l) {
            const bool last = l == np.L - 1;
            const float* X1 = last ? rows + P.row_seed * FS : netrows + P.s_row[l + 1] * FS;
            const float* Z1 = l == 0 ? r_in : netrows + P.a_row[l] * FS;
            const float* X2 = l == 0 ? nullptr : (last ? rows + P.row_one * FS : netrows + P.g_row[l + 1] * FS);
            const float* Z2 = l == 0 ? nullptr : netrows + P.t_row[l] * FS;
            rot = outer_layer(rot, tid, nt, np.dims[l + 1], np.dims[l], X1, Z1, X2, Z2, FS, F, pn + np.gw_off[l], pn + np.gb_off[l]);
          }
        }
        __syncthreads();
        PT_MARK(8);
      }
    }  // networks

    if (!GRAD) {
      // ---- batch sums of this tile in fp64: S0, S1[i], S2[i][j], SD[i]
      const int ns = P.n_stats;
      for (int f = tid; f < F; f += nt) {
        const long long fr = f_base + f;
        if (fr < B)
          for (int i = 0; i < k; ++i) y_io[(size_t)i * B + fr] = rows[(P.row_y + i) * FS + f];
      }
      const int lane = tid & 31, warp = tid >> 5, nwarp_f = F / 32;
      if (tid < F) {
        const double wf = rows[P.row_w * FS + tid];
        for (int s = 0; s < ns; ++s) {
          double v;
          if (s == 0) v = wf;
          else if (s < 1 + k) v = wf * rows[(P.row_y + s - 1) * FS + tid];
          else if (s < 1 + k + k * k) {
            const int i = (s - 1 - k) / k, j = (s - 1 - k) % k;
            v = wf * (double)rows[(P.row_y + i) * FS + tid] * (double)rows[(P.row_y + j) * FS + tid];
          } else v = wf * rows[(P.row_D + s - 1 - k - k * k) * FS + tid];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) red[s * 4 + warp] = v;
        }
      }
      __syncthreads();
      if (tid < ns) {
        double v = 0.0;
        for (int q = 0; q < nwarp_f; ++q) v += red[tid * 4 + q];
        stat_acc += v;
      }
      __syncthreads();
      PT_MARK(9);
    }
  }  // tiles
  PT_FLUSH;
  if (!GRAD && tid < P.n_stats) part[tid] = stat_acc;
}

struct EigW {
  double v[kMaxK];
};

// loss, eigenvalues, ordering, objective, penalty and the pass-2 coefficients (core.py:406-410,426-455).
__global__ void eigen_combine_kernel(const double* __restrict__ S, int k, double alpha, double beta, int sort,
                                     const EigW eig_w, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double S0 = S[0];
  const double* S1 = S + 1;
  const double* S2 = S + 1 + k;
  const double* SD = S + 1 + k + k * k;
  double mean[kMaxK], var[kMaxK], E[kMaxK], eig[kMaxK], omega[kMaxK];
  int cvec[kMaxK];
  for (int i = 0; i < k; ++i) {
    mean[i] = S1[i] / S0;
    var[i] = S2[i * k + i] / S0 - mean[i] * mean[i];
    E[i] = SD[i] / (beta * S0);
    eig[i] = E[i] / var[i];
    cvec[i] = i;
  }
  if (sort) {   // ascending, stable (np.argsort, core.py:432)
    for (int i = 1; i < k; ++i) {
      const int c = cvec[i];
      int j = i - 1;
      while (j >= 0 && eig[cvec[j]] > eig[c]) cvec[j + 1] = cvec[j], --j;
      cvec[j + 1] = c;
    }
  }
  for (int r = 0; r < k; ++r) omega[cvec[r]] = eig_w.v[r];
  double obj = 0.0, pen = 0.0;
  for (int i = 0; i < k; ++i) {
    obj += omega[i] * eig[i];
    pen += (var[i] - 1.0) * (var[i] - 1.0);
  }
  double* C2 = out + 3 + 4 * k;
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      if (i == j) {
        const double cv = -omega[i] * E[i] / (var[i] * var[i]) + 2.0 * alpha * (var[i] - 1.0);
        C2[i * k + j] = 2.0 * cv / S0;
      } else {
        const double cov = S2[i * k + j] / S0 - mean[i] * mean[j];
        if (j > i) pen += cov * cov;
        C2[i * k + j] = 2.0 * alpha * cov / S0;
      }
    }
  out[0] = obj + alpha * pen;
  out[1] = obj;
  out[2] = pen;
  for (int r = 0; r < k; ++r) {
    out[3 + r] = eig[cvec[r]];
    out[3 + k + r] = (double)cvec[r];
    out[3 + 2 * k + r] = mean[r];
    out[3 + 3 * k + r] = omega[r] / (beta * S0 * var[r]);
    out[3 + 4 * k + k * k + r] = 0.0;   // a0
  }
}

static int eigen_launch(bool grad, const float* x, const float* w, int64_t B, const cvf_preproc* pp, const cvf_mlp* net,
                        int k, const float* params, float* y_io, const double* combine, const float* seed_extra, double* out,
                        void* workspace, size_t ws_bytes, int scratch_valid, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EigenPlan P;
  int e = build_plan(pp, net, k, &P);
  if (e) return e;
  if (!x || !w || !params || !y_io || !out || !workspace || B < 1 || (grad && !combine)) {
    set_error("null pointer or empty batch");
    return CVF_E_ARG;
  }
  if (g_eigen_path == 0 && fast_eigen_supported(pp, P.net, k)) {
    if (grad)
      return fast_eigen_grad(pp, P.net, k, x, w, B, params, combine, seed_extra, out, workspace, ws_bytes, scratch_valid, stream);
    return fast_eigen_stats(pp, P.net, k, x, w, B, params, y_io, out, workspace, ws_bytes, stream);
  }
  e = finish_plan(&P, pp->kind == 1 && pp->positions_only != 0);   // largest tile that fits
  if (e) return e;
  const int n_part = grad ? k * P.net.n_params : P.n_stats;
  const long long n_tiles = (B + P.F - 1) / P.F;
  int grid = sm_count();
  if (n_tiles < grid) grid = (int)n_tiles;
  if ((size_t)grid * n_part * sizeof(double) > ws_bytes) {
    set_error("workspace too small: %zu < %zu", ws_bytes, (size_t)grid * n_part * sizeof(double));
    return CVF_E_WORKSPACE;
  }
  if (grad) {
    CVF_CUDA(cudaFuncSetAttribute(eigen_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes));
    CVF_LAUNCH(K_EIGEN_GRAD, stream, eigen_kernel<true, 2><<<grid, P.nthreads, P.smem_bytes, stream>>>(P, x, w, B, params, y_io, combine, (double*)workspace, seed_extra));
  } else {
    CVF_CUDA(cudaFuncSetAttribute(eigen_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes));
    CVF_LAUNCH(K_EIGEN_STATS, stream, eigen_kernel<false, 2><<<grid, P.nthreads, P.smem_bytes, stream>>>(P, x, w, B, params, y_io, combine, (double*)workspace, nullptr));
  }
  CVF_CUDA(cudaGetLastError());
  CVF_LAUNCH(K_REDUCE, stream, reduce_partials_kernel<<<(n_part + 127) / 128, 128, 0, stream>>>((const double*)workspace, grid, n_part, 0, n_part, out));
  CVF_CUDA(cudaGetLastError());
  return 0;
}

// ---- transfer-operator branch (core.py:412-416,428,440): forward-only on X and on the time-lagged X' --------------
// mode 0: sx[i] = sum_f w_f (y'_i - y_i)^2 (per-block partials);  mode 1: extra[i][f] = E_i w_f (y_i - y'_i)
__global__ void __launch_bounds__(256)
tlag_terms_kernel(const float* __restrict__ y, const float* __restrict__ yl, const float* __restrict__ w, long long B, int k,
                  const double* __restrict__ coef, double* __restrict__ part, float* __restrict__ extra) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = 0; i < k; ++i) {
    double acc = 0.0;
    const float e = extra ? (float)coef[i] : 0.0f;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (long long)gridDim.x * blockDim.x) {
      const float d = yl[(size_t)i * B + f] - y[(size_t)i * B + f];
      if (extra) extra[(size_t)i * B + f] = -e * w[f] * d;
      else acc += (double)w[f] * (double)d * (double)d;
    }
    if (extra) continue;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int q = 0; q < 8; ++q) t += red[q];
      part[(size_t)blockIdx.x * k + i] = t;
    }
    __syncthreads();
  }
}

// Loss, eigenvalues, ordering and the coefficient vectors of both backward passes.  Stats layout as cvf_eigen_stats
// (S0, S1[k], S2[k*k], SD[k]; SD unused).  Outputs in the cvf_eigen_combine layout with cD = 0, plus E[k] at the end of `out`.
__global__ void tlag_combine_kernel(const double* __restrict__ S, const double* __restrict__ Sl, const double* __restrict__ SX, int k,
                                    double alpha, double tau, int sort, const EigW eig_w, double* __restrict__ out,
                                    double* __restrict__ out_lag) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double S0 = S[0], S0l = Sl[0];
  const double *S1 = S + 1, *S2 = S + 1 + k, *S1l = Sl + 1, *S2l = Sl + 1 + k;
  double mean[kMaxK], var[kMaxK], meanl[kMaxK], varl[kMaxK], eig[kMaxK], a[kMaxK], b[kMaxK];
  int cvec[kMaxK];
  for (int i = 0; i < k; ++i) {
    mean[i] = S1[i] / S0, var[i] = S2[i * k + i] / S0 - mean[i] * mean[i];
    meanl[i] = S1l[i] / S0l, varl[i] = S2l[i * k + i] / S0l - meanl[i] * meanl[i];
    eig[i] = SX[i] / (tau * S0 * (var[i] + varl[i]));   // core.py:428
    cvec[i] = i, b[i] = 0.0;
  }
  if (sort) {
    for (int i = 1; i < k; ++i) {
      const int c = cvec[i];
      int j = i - 1;
      while (j >= 0 && eig[cvec[j]] > eig[c]) cvec[j + 1] = cvec[j], --j;
      cvec[j + 1] = c;
    }
  }
  // core.py:440 -- the numerator of term idx is that of network idx, its denominator that of network cvec[idx]
  double obj = 0.0, pen = 0.0;
  for (int idx = 0; idx < k; ++idx) {
    const int c = cvec[idx];
    const double V = var[c] + varl[c];
    obj += eig_w.v[idx] * SX[idx] / (tau * S0 * V);
    a[idx] = eig_w.v[idx] / (tau * S0 * V);
    b[c] -= eig_w.v[idx] * SX[idx] / (tau * S0 * V * V);
  }
  double* C2 = out + 3 + 4 * k;
  double* C2l = out_lag + 3 + 4 * k;
  for (int i = 0; i < k; ++i) {
    pen += (var[i] - 1.0) * (var[i] - 1.0);
    for (int j = 0; j < k; ++j) {
      C2l[i * k + j] = i == j ? 2.0 * b[i] / S0l : 0.0;
      if (i == j) {
        C2[i * k + j] = 2.0 * (b[i] + 2.0 * alpha * (var[i] - 1.0)) / S0;
      } else {
        const double cov = S2[i * k + j] / S0 - mean[i] * mean[j];
        if (j > i) pen += cov * cov;
        C2[i * k + j] = 2.0 * alpha * cov / S0;
      }
    }
  }
  out[0] = obj + alpha * pen, out[1] = obj, out[2] = pen;
  for (int r = 0; r < k; ++r) {
    out[3 + r] = eig[cvec[r]], out[3 + k + r] = (double)cvec[r];
    out[3 + 2 * k + r] = mean[r], out[3 + 3 * k + r] = 0.0;
    out_lag[3 + r] = eig[cvec[r]], out_lag[3 + k + r] = (double)cvec[r];
    out_lag[3 + 2 * k + r] = meanl[r], out_lag[3 + 3 * k + r] = 0.0;
    out[3 + 4 * k + k * k + r] = 0.0, out_lag[3 + 4 * k + k * k + r] = 0.0;   // a0
    out[3 + 5 * k + k * k + r] = 2.0 * a[r];   // E
  }
  out_lag[0] = out[0], out_lag[1] = out[1], out_lag[2] = out[2];
}

}  // namespace cvf

using namespace cvf;

#ifdef CVF_PHASE_TIMERS
extern "C" int cvf_debug_phase_cycles(unsigned long long* out16, int reset) {
  cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" int32_t cvf_eigen_num_stats(int32_t k) { return 1 + 2 * k + k * k; }
extern "C" int32_t cvf_eigen_num_combine(int32_t k) { return 3 + 5 * k + k * k; }

extern "C" size_t cvf_eigen_workspace_bytes(const cvf_preproc* pp, const cvf_mlp* net, int32_t k, int64_t B) {
  NetPlan np;
  if (!pp || make_net_plan(net, &np) || k < 1 || k > kMaxK || B < 1) return 0;
  size_t n = (size_t)k * np.n_params;
  const size_t ns = (size_t)cvf_eigen_num_stats(k);
  if (ns > n) n = ns;
  size_t bytes = n * sizeof(double) * (size_t)sm_count();
  if (fast_eigen_supported(pp, np, k)) {
    const size_t fb = fast_eigen_workspace_bytes(pp, np, k, B);
    if (fb > bytes) bytes = fb;
  }
  return bytes;
}

extern "C" int cvf_eigen_set_path(int32_t mode) {
  if (mode != 0 && mode != 1) {
    set_error("cvf_eigen_set_path: mode must be 0 (auto) or 1 (general kernels)");
    return CVF_E_ARG;
  }
  g_eigen_path = mode;
  return 0;
}

extern "C" int cvf_eigen_path(const cvf_preproc* pp, const cvf_mlp* net, int32_t k) {
  NetPlan np;
  if (!pp || make_net_plan(net, &np)) return CVF_E_ARG;
  return g_eigen_path == 0 && fast_eigen_supported(pp, np, k) ? 1 : 0;
}

extern "C" int cvf_eigen_stats(const float* x, const float* w, int64_t B, const cvf_preproc* pp, const cvf_mlp* net, int32_t k,
                               const float* params, float* y_out, double* stats_out, void* workspace, size_t workspace_bytes,
                               void* stream) {
  return eigen_launch(false, x, w, B, pp, net, k, params, y_out, nullptr, nullptr, stats_out, workspace, workspace_bytes, 0, stream);
}

extern "C" int cvf_eigen_combine(const double* stats, int32_t k, double alpha, const double* eig_w, double beta, int32_t sort,
                                 double* combine_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!stats || !eig_w || !combine_out || k < 1 || k > kMaxK) {
    set_error("cvf_eigen_combine: bad argument");
    return CVF_E_ARG;
  }
  EigW ew;   // eig_w is a HOST array: it travels by value as a kernel parameter
  for (int i = 0; i < kMaxK; ++i) ew.v[i] = i < k ? eig_w[i] : 0.0;
  CVF_LAUNCH(K_EIGEN_COMBINE, stream, eigen_combine_kernel<<<1, 32, 0, stream>>>(stats, k, alpha, beta, sort, ew, combine_out));
  CVF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cvf_eigen_grad(const float* x, const float* w, int64_t B, const cvf_preproc* pp, const cvf_mlp* net, int32_t k,
                              const float* params, const float* y_in, const double* combine, const float* seed_extra,
                              double* grad_out, void* workspace, size_t workspace_bytes, int32_t scratch_valid, void* stream) {
  return eigen_launch(true, x, w, B, pp, net, k, params, const_cast<float*>(y_in), combine, seed_extra, grad_out, workspace,
                      workspace_bytes, scratch_valid, stream);
}

extern "C" int cvf_eigen_tlag_terms(const float* y, const float* y_lag, const float* w, int64_t B, int32_t k, const double* coef,
                                    double* sx_out, float* extra_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!y || !y_lag || !w || B < 1 || k < 1 || k > kMaxK || (!extra_out && (!sx_out || !workspace)) || (extra_out && !coef)) {
    set_error("cvf_eigen_tlag_terms: bad argument");
    return CVF_E_ARG;
  }
  long long grid = (long long)sm_count() * 4;
  if ((B + 255) / 256 < grid) grid = (B + 255) / 256;
  if (!extra_out && (size_t)grid * k * sizeof(double) > workspace_bytes) {
    set_error("workspace too small");
    return CVF_E_WORKSPACE;
  }
  CVF_LAUNCH(K_EIGEN_STATS, stream, tlag_terms_kernel<<<(int)grid, 256, 0, stream>>>(y, y_lag, w, B, k, coef, (double*)workspace, extra_out));
  CVF_CUDA(cudaGetLastError());
  if (!extra_out) {
    CVF_LAUNCH(K_REDUCE, stream, reduce_partials_kernel<<<1, 32, 0, stream>>>((const double*)workspace, (int)grid, k, 0, k, sx_out));
    CVF_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int cvf_eigen_tlag_combine(const double* stats, const double* stats_lag, const double* sx, int32_t k, double alpha,
                                      const double* eig_w, double tau, int32_t sort, double* combine_out, double* combine_lag_out,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!stats || !stats_lag || !sx || !eig_w || !combine_out || !combine_lag_out || k < 1 || k > kMaxK || !(tau > 0.0)) {
    set_error("cvf_eigen_tlag_combine: bad argument");
    return CVF_E_ARG;
  }
  EigW ew;
  for (int i = 0; i < kMaxK; ++i) ew.v[i] = i < k ? eig_w[i] : 0.0;
  CVF_LAUNCH(K_EIGEN_COMBINE, stream,
             tlag_combine_kernel<<<1, 32, 0, stream>>>(stats, stats_lag, sx, k, alpha, tau, sort, ew, combine_out, combine_lag_out));
  CVF_CUDA(cudaGetLastError());
  return 0;
}
