// 2-D Euler on curvilinear structured quadrilaterals, the two-kernel form: a face kernel (common fluxes over the face
// connectivity) and an element kernel (derivative + correction + RK stage) per stage, lanes along i so that every
// state / metric plane is read in 256-byte runs.  The arithmetic lives in frb_euler2d_curv_elem.cuh
// (reference: dev/parallelogram.jl:80-165, dev/cylinder2.jl:52-164); measurements in profiles/r01_curv.md.
//
// Algorithmic bytes per DOF-update: the state terms of the rectangular path (16 B / 24 B) plus the
// metric, which no longer is two scalars: 4 doubles of iJ per solution point shared by the 4 variables
// = 8 B per DOF (+ normals and flux-point factors, O(1/nsp) of that).
#include "frb_euler2d_curv_elem.cuh"

namespace {

// thread = (element i = lane, flux point p = threadIdx.y) computes the x face left of the element and the y face
// below it.  The NSP point threads of a block read every value of their elements twice (as a row for the x trace,
// as a column for the y trace): one DRAM pass, the second read is an L1 hit; the left neighbour is the next lane,
// the lower neighbour's block ran just before (L2).
// IX = unsigned when every array of the problem stays below 2^32 elements (the launcher checks): both kernels are
// bound by instruction issue, and 64-bit index arithmetic was a third of their instructions.  FLUX: the common
// flux as a compile-time choice (branch-free forms only, no IEEE-division slow paths in the instruction stream).
template <int NSP, typename IX, int FLUX>
__global__ void __launch_bounds__(32 * NSP)
euler2d_curv_face_kernel(const double *__restrict__ u, double *__restrict__ fx, double *__restrict__ fy,
                         CurvGeom g, double gamma, FrbOps ops) {
  const int i = blockIdx.x * 32 + threadIdx.x + 1;
  const int j = blockIdx.y + 1;
  const int p = threadIdx.y;
  if (i > g.nx + 1) return;
  frbcurv::face_xy<NSP, IX, FLUX>(i, j, p, j <= g.ny, i <= g.nx, u, fx, fy, g, gamma, ops);
}

// thread = (element i = lane, point row l = threadIdx.y); one block = 32 consecutive elements of row j.
// tile holds f2 = (iJ [F; G])[2] of the block's elements, [l][k][m][lane], fyt the y common fluxes of their
// bottom / top faces, [side][p][m][lane]; both conflict-free (lane fastest).
// 3 blocks / SM (168 registers at p3) measured faster than 4 (128 registers, spills): profiles/r01_curv.md.
// The correction's flux traces are folded into the derivative matrix (row_xpass / row_ypass with FOLD).
template <int NSP, typename IX>
__global__ void __launch_bounds__(32 * NSP, 3)
euler2d_curv_elem_kernel(const double *__restrict__ u, const double *__restrict__ ua,
                         const double *__restrict__ fx, const double *__restrict__ fy, double *__restrict__ out,
                         CurvGeom g, double gamma, FrbOps ops, FrbStage st) {
  __shared__ double tile[NSP * NSP * 4 * 32];
  __shared__ double fyt[2 * NSP * 4 * 32];
  const int lane = threadIdx.x, l = threadIdx.y;
  const int i = blockIdx.x * 32 + lane + 1;
  const int j = blockIdx.y + 1;
  const bool active = i <= g.nx;
  frbcurv::RowCarry<NSP> c;
  if (active) frbcurv::row_xpass<NSP, IX, true>(i, j, l, u, fx, fy, g, gamma, ops, tile + lane, fyt + lane, 32, c);
  __syncthreads();
  if (active) frbcurv::row_ypass<NSP, IX, true>(i, j, l, ua, out, g, ops, st, tile + lane, fyt + lane, 32, c);
}

// Per-step boundary fill of dev/cylinder2.jl:176-187 on the ring-embedded array (interior nx = nr - 1,
// column nx+1 = the script's cell nr).  theta ghosts: u[i, 0, k, l, :] = flip_y(u[i, 1, nsp+1-k, nsp+1-l, :]),
// u[i, ny+1, k, l, :] = flip_y(u[i, ny, nsp+1-k, nsp+1-l, :]) for i = 1..nx+1 (the script's "4 - k" at deg 2).
__global__ void ghost_cyl_theta_kernel(double *__restrict__ u, int nx, int ny, int nsp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int p = blockIdx.y;  // k + nsp (l + nsp m)
  if (i > nx + 1) return;
  const int k = p % nsp, l = (p / nsp) % nsp, m = p / (nsp * nsp);
  const int ps = (nsp - 1 - k) + nsp * ((nsp - 1 - l) + nsp * m);
  const size_t NXG = nx + 2, NE = NXG * (size_t)(ny + 2);
  const double sg = m == 2 ? -1.0 : 1.0;
  u[i + NE * p] = sg * u[i + NXG * 1 + NE * ps];
  u[i + NXG * (size_t)(ny + 1) + NE * p] = sg * u[i + NXG * (size_t)ny + NE * ps];
}
// outer column: u[nx+1, j, :] = u[nx, j, :] for j = 1..ny/2
__global__ void ghost_cyl_outer_kernel(double *__restrict__ u, int nx, int ny) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int p = blockIdx.y;
  if (j > ny / 2) return;
  const size_t NXG = nx + 2, NE = NXG * (size_t)(ny + 2);
  u[nx + 1 + NXG * j + NE * p] = u[nx + NXG * j + NE * p];
}

}  // namespace

int frb_launch_ghost_cylinder(frb_prob_t p, double *u) {
  const int nplanes = 4 * p->nsp * p->nsp;
  cudaStream_t s = p->ctx->stream;
  dim3 blk(128), g1((p->nx + 1 + 127) / 128, nplanes), g2((p->ny / 2 + 127) / 128, nplanes);
  ghost_cyl_theta_kernel<<<g1, blk, 0, s>>>(u, p->nx, p->ny, p->nsp);
  int n = 1;
  if (p->ny / 2 > 0) {
    ghost_cyl_outer_kernel<<<g2, blk, 0, s>>>(u, p->nx, p->ny);
    ++n;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "ghost_cyl kernels", __FILE__, __LINE__);
  return n;
}

int frb_launch_euler2d_curv(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  if (u == out) {  // neighbours' blocks are read: never in place
    frb_set_error("euler2d_curv: the stage cannot run in place");
    return FRB_ERR_STATE;
  }
  if (p->ny + 1 > 65535) {
    frb_set_error("euler2d_curv: ny must be < 65535");
    return FRB_ERR_ARG;
  }
  CurvGeom g;
  g.nx = p->nx; g.ny = p->ny;
  g.iJ = p->curv_iJ; g.n1 = p->curv_n1; g.n2 = p->curv_n2; g.fpc = p->curv_fpc;
  g.fy_row = (p->curv_flags & FRB_CURV_FY_ROW_INDEX) ? 1 : 0;
  g.wall_xlo = (p->curv_flags & FRB_CURV_WALL_XLO) ? 1 : 0;
  g.flux = p->flux;
  g.vert = p->curv_vert;
  for (int q = 0; q < FRB_NSPMAX; ++q) g.r[q] = p->curv_r[q];
  if (st.nested) { st.cdt *= st.cb; st.nested = 0; }
  if (st.rhs_only) { st.ca = 0.0; st.cb = 0.0; st.cdt = 1.0; st.use_a = 0; }  // out = L(u), branch-free in the kernel
  if (!st.use_a) st.ca = 0.0;
  // FRB_KERNEL_CURV_MARCH (or FRB_CURV_MARCH=1 in the environment, read per launch: A / B in one process): one
  // marching launch (frb_euler2d_curv_march.cu) -- half the DRAM traffic, but slower than the two launches below at
  // the register budget a row owner needs (profiles/r02_summary.md, section F)
  if (p->kernel_kind == FRB_KERNEL_CURV_MARCH || getenv("FRB_CURV_MARCH") != nullptr) {
    const bool idx32 = (double)(p->nx + 2) * (p->ny + 2) * p->nsp * p->nsp * 4 < 4294967296.0;  // its index width
    if (!g.vert && idx32) return frb_launch_euler2d_curv_march(p, u, ua, out, g, st);
    if (p->kernel_kind == FRB_KERNEL_CURV_MARCH) {
      frb_set_error("euler2d_curv: the marching kernel needs the stored metric and arrays below 2^32 elements");
      return FRB_ERR_STATE;
    }
  }
  // default: one launch per stage, every block evaluates the fluxes of its own faces (frb_euler2d_curv_fused.cu);
  // FRB_KERNEL_GENERIC (or FRB_CURV_TWO_KERNELS=1, read per launch) and the vertex metric take the two launches below
  if (p->kernel_kind != FRB_KERNEL_GENERIC && !g.vert && getenv("FRB_CURV_TWO_KERNELS") == nullptr && p->ny <= 65535)
    return frb_launch_euler2d_curv_fused(p, u, ua, out, g, st);
  const size_t nfx = (size_t)(p->nx + 1) * p->ny * p->nsp * 4, nfy = (size_t)p->nx * (p->ny + 1) * p->nsp * 4;
  if (!p->curv_flux) FRB_CUDA(cudaMalloc(&p->curv_flux, sizeof(double) * (nfx + nfy)));
  double *fx = p->curv_flux, *fy = p->curv_flux + nfx;
  cudaStream_t s = p->ctx->stream;
  dim3 fb(32, p->nsp), fg((p->nx + 1 + 31) / 32, p->ny + 1);
  dim3 eb(32, p->nsp), eg((p->nx + 31) / 32, p->ny);
  // every array below 2^32 elements (state / metric: NE nsp^2 4; the flux and normal tables are smaller)
  const bool ix32 = (double)(p->nx + 2) * (p->ny + 2) * p->nsp * p->nsp * 4 < 4294967296.0 &&
                    getenv("FRB_CURV_IX64") == nullptr;
  const int key = p->nsp * 100 + (ix32 ? 10 : 0) + g.flux;
  switch (key) {
#define FRB_CURV_CASE(N, IX, I, F)                                                                        \
  case N * 100 + I * 10 + F:                                                                              \
    euler2d_curv_face_kernel<N, IX, F><<<fg, fb, 0, s>>>(u, fx, fy, g, p->gamma, p->ops);                  \
    euler2d_curv_elem_kernel<N, IX><<<eg, eb, 0, s>>>(u, ua, fx, fy, out, g, p->gamma, p->ops, st);        \
    break;
#define FRB_CURV_CASES(N)                                                                                 \
  FRB_CURV_CASE(N, unsigned, 1, 0) FRB_CURV_CASE(N, unsigned, 1, 1) FRB_CURV_CASE(N, unsigned, 1, 2)     \
  FRB_CURV_CASE(N, size_t, 0, 0) FRB_CURV_CASE(N, size_t, 0, 1) FRB_CURV_CASE(N, size_t, 0, 2)
    FRB_CURV_CASES(2)
    FRB_CURV_CASES(3)
    FRB_CURV_CASES(4)
#undef FRB_CURV_CASES
#undef FRB_CURV_CASE
    default: frb_set_error("euler2d_curv: deg must be in 1..3, flux HLL / LF / ROE"); return FRB_ERR_ARG;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "euler2d_curv kernels", __FILE__, __LINE__);
  return 2;
}
