// Row-chunk layout (frb_rc.cuh): conversion to / from the reference memory image and the small
// per-step utilities of the time loop restated for it -- periodic / copy ghost fill
// (example/euler2d_wave.jl:127-132, :159-164; shock-vortex.jl:324-326), the copy of the frozen
// ghost ring into the stage buffers, and the slab-boundary row push of the multi-GPU path.
#include "frb_internal.cuh"
#include "frb_rc.cuh"

namespace {

// one thread per RC slot: gather from / scatter to the reference image
__global__ void rc_from_ref_kernel(const double *__restrict__ ref, double *__restrict__ rc, RcGeom g) {
  const int lane = threadIdx.x & 31;
  const int plane = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int s = blockIdx.y, j = blockIdx.z;
  if (plane >= g.nplanes) return;
  const int i = kRcOwn * s + lane;
  const size_t NXG = g.nx + 2, NE = NXG * (size_t)(g.ny + 2);
  double v = 0.0;  // padding lanes of the last strip
  if (i <= g.nx + 1) v = ref[i + NXG * j + NE * plane];
  rc[rc_index(g, j, s, plane, lane)] = v;
}

// interior: only the owned elements (du of f!: the ghosts of du stay 0)
__global__ void rc_to_ref_kernel(const double *__restrict__ rc, double *__restrict__ ref, RcGeom g, int interior) {
  const int lane = threadIdx.x & 31;
  const int plane = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int s = blockIdx.y, j = blockIdx.z;
  if (plane >= g.nplanes) return;
  const int i = kRcOwn * s + lane;
  if (i > g.nx + 1) return;
  if (interior && (i == 0 || i == g.nx + 1 || j == 0 || j == g.ny + 1)) return;
  int sp, lp;
  rc_primary(g, i, &sp, &lp);
  if (sp != s) return;  // a duplicate: the primary copy is written by its own strip
  const size_t NXG = g.nx + 2, NE = NXG * (size_t)(g.ny + 2);
  ref[i + NXG * j + NE * plane] = rc[rc_index(g, j, s, plane, lane)];
}

__device__ __forceinline__ double rc_load_col(const double *u, const RcGeom &g, int i, int j, int plane) {
  int s, lane;
  rc_primary(g, i, &s, &lane);
  return u[rc_index(g, j, s, plane, lane)];
}
__device__ __forceinline__ void rc_store_col(double *u, const RcGeom &g, int i, int j, int plane, double v) {
  int s, lane, s2, l2;
  rc_primary(g, i, &s, &lane);
  u[rc_index(g, j, s, plane, lane)] = v;
  if (rc_duplicate(g, s, lane, &s2, &l2)) u[rc_index(g, j, s2, plane, l2)] = v;
}

// u[0,j,p] = sg * u[srcL,j,p] (if do_left), u[nx+1,j,p] = sg * u[srcR,j,p]   for all rows, planes
__global__ void rc_ghost_x_kernel(double *__restrict__ u, RcGeom g, int npp, int srcL, int srcR,
                                  int flip_var, int do_left) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (j > g.ny + 1) return;
  const double sg = (p / npp == flip_var) ? -1.0 : 1.0;
  if (do_left) rc_store_col(u, g, 0, j, p, sg * rc_load_col(u, g, srcL, j, p));
  rc_store_col(u, g, g.nx + 1, j, p, sg * rc_load_col(u, g, srcR, j, p));
}

// row 0 = sg * row srcB, row ny+1 = sg * row srcT (whole RC rows: every copy of every column)
__global__ void rc_ghost_y_kernel(double *__restrict__ u, RcGeom g, int npp, int srcB, int srcT, int flip_var) {
  const size_t off = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (off >= g.row) return;
  const int plane = (int)(off % g.chunk) >> 5;
  const double sg = (plane / npp == flip_var) ? -1.0 : 1.0;
  u[off] = sg * u[off + g.row * (size_t)srcB];
  u[off + g.row * (size_t)(g.ny + 1)] = sg * u[off + g.row * (size_t)srcT];
}

// ghost rows (optional) and ghost columns of src -> dst
__global__ void rc_ring_rows_kernel(const double *__restrict__ src, double *__restrict__ dst, RcGeom g,
                                    int row0, int rowN) {
  const size_t off = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (off >= g.row) return;
  if (row0) dst[off] = src[off];
  const size_t top = g.row * (size_t)(g.ny + 1);
  if (rowN) dst[off + top] = src[off + top];
}
__global__ void rc_ring_cols_kernel(const double *__restrict__ src, double *__restrict__ dst, RcGeom g) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int p = blockIdx.y;
  if (j > g.ny) return;
  rc_store_col(dst, g, 0, j, p, rc_load_col(src, g, 0, j, p));
  rc_store_col(dst, g, g.nx + 1, j, p, rc_load_col(src, g, g.nx + 1, j, p));
}

// my first owned row -> the upper halo row of the rank below; my last owned row -> the lower
// halo row of the rank above (direct stores into peer memory)
__global__ void rc_row_push_kernel(const double *__restrict__ src, double *__restrict__ dst_lo,
                                   double *__restrict__ dst_hi, RcGeom g, int nyl_lo, int npp, int flip_var) {
  const size_t off = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (off >= g.row) return;
  const int plane = (int)(off % g.chunk) >> 5;
  const double sg = (plane / npp == flip_var) ? -1.0 : 1.0;
  if (dst_lo) dst_lo[off + g.row * (size_t)(nyl_lo + 1)] = sg * src[off + g.row];
  if (dst_hi) dst_hi[off] = sg * src[off + g.row * (size_t)g.ny];
}

// slab-parallel per-step work in one launch each (strong scaling: a step is a few hundred microseconds, every
// small launch counts).  (1) the x half of the ghost fill on u_n AND the copy of the frozen ghost columns into the
// two stage buffers (rc_ghost_x_kernel + 2 x rc_ring_cols_kernel);
__global__ void rc_ghost_x_ring_kernel(double *__restrict__ u, double *__restrict__ d1, double *__restrict__ d2,
                                       RcGeom g, int npp, int srcL, int srcR, int flip_var) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (j > g.ny + 1) return;
  const double sg = (p / npp == flip_var) ? -1.0 : 1.0;
  const double vl = sg * rc_load_col(u, g, srcL, j, p), vr = sg * rc_load_col(u, g, srcR, j, p);
  rc_store_col(u, g, 0, j, p, vl);
  rc_store_col(u, g, g.nx + 1, j, p, vr);
  if (j >= 1 && j <= g.ny) {
    if (d1) { rc_store_col(d1, g, 0, j, p, vl); rc_store_col(d1, g, g.nx + 1, j, p, vr); }
    if (d2) { rc_store_col(d2, g, 0, j, p, vl); rc_store_col(d2, g, g.nx + 1, j, p, vr); }
  }
}
// (2) the frozen seam rows of the step into all three buffers of the seam neighbour (3 x rc_row_push_kernel)
struct RcPush3 {
  double *lo[3], *hi[3];
};
__global__ void rc_row_push3_kernel(const double *__restrict__ src, RcPush3 d, RcGeom g, int nyl_lo, int npp,
                                    int flip_var) {
  const size_t off = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (off >= g.row) return;
  const int plane = (int)(off % g.chunk) >> 5;
  const double sg = (plane / npp == flip_var) ? -1.0 : 1.0;
  const double v1 = sg * src[off + g.row], vn = sg * src[off + g.row * (size_t)g.ny];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (d.lo[r]) d.lo[r][off + g.row * (size_t)(nyl_lo + 1)] = v1;
    if (d.hi[r]) d.hi[r][off] = vn;
  }
}

int check_launch_rc(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, what, __FILE__, __LINE__);
  return 0;
}

}  // namespace

static RcGeom geom_of(frb_prob_t p) { return rc_geom(p->nx, p->ny, p->nsp); }

int frb_rc_from_ref(frb_prob_t p, const double *ref, double *rc) {
  const RcGeom g = geom_of(p);
  dim3 blk(256), grd((g.nplanes + 7) / 8, g.ns, g.ny + 2);
  rc_from_ref_kernel<<<grd, blk, 0, p->ctx->stream>>>(ref, rc, g);
  if (int r = check_launch_rc("rc_from_ref_kernel")) return r;
  return 1;
}

int frb_rc_to_ref(frb_prob_t p, const double *rc, double *ref, bool interior_only) {
  const RcGeom g = geom_of(p);
  dim3 blk(256), grd((g.nplanes + 7) / 8, g.ns, g.ny + 2);
  rc_to_ref_kernel<<<grd, blk, 0, p->ctx->stream>>>(rc, ref, g, interior_only ? 1 : 0);
  if (int r = check_launch_rc("rc_to_ref_kernel")) return r;
  return 1;
}

static int rc_ghost_x(frb_prob_t p, double *u, const RcGeom &g, int srcL, int srcR, int flip, int do_left) {
  dim3 blk(128), grd((g.ny + 2 + 127) / 128, g.nplanes);
  rc_ghost_x_kernel<<<grd, blk, 0, p->ctx->stream>>>(u, g, p->nsp * p->nsp, srcL, srcR, flip, do_left);
  return check_launch_rc("rc_ghost_x_kernel");
}
static int rc_ghost_y(frb_prob_t p, double *u, const RcGeom &g, int srcB, int srcT, int flip) {
  dim3 blk(256), grd((unsigned)((g.row + 255) / 256));
  rc_ghost_y_kernel<<<grd, blk, 0, p->ctx->stream>>>(u, g, p->nsp * p->nsp, srcB, srcT, flip);
  return check_launch_rc("rc_ghost_y_kernel");
}

// same modes, same pass order as frb_launch_ghost_fill2d
int frb_rc_ghost_fill(frb_prob_t p, double *u, int mode) {
  const RcGeom g = geom_of(p);
  int r;
  if (mode == FRB_GHOST_WAVE_X) {
    if ((r = rc_ghost_x(p, u, g, p->nx, 1, -1, 1))) return r;
    if ((r = rc_ghost_y(p, u, g, p->ny, 1, 2))) return r;
  } else if (mode == FRB_GHOST_WAVE_Y) {
    if ((r = rc_ghost_y(p, u, g, p->ny, 1, -1))) return r;
    if ((r = rc_ghost_x(p, u, g, p->nx, 1, 1, 1))) return r;
  } else if (mode == FRB_GHOST_COPY) {
    if ((r = rc_ghost_y(p, u, g, 1, p->ny, -1))) return r;
    if ((r = rc_ghost_x(p, u, g, 0, p->nx, -1, 0))) return r;
  } else if (mode == FRB_GHOST_PERIODIC) {
    if ((r = rc_ghost_x(p, u, g, p->nx, 1, -1, 1))) return r;
    if ((r = rc_ghost_y(p, u, g, p->ny, 1, -1))) return r;
  } else {
    frb_set_error("unknown ghost mode");
    return FRB_ERR_ARG;
  }
  return 2;
}

// the x half only (slab-parallel path: the y half is an exchange with the neighbouring ranks)
int frb_rc_ghost_x(frb_prob_t p, double *u, int mode) {
  const RcGeom g = geom_of(p);
  int r;
  if (mode == FRB_GHOST_WAVE_X) r = rc_ghost_x(p, u, g, p->nx, 1, -1, 1);
  else if (mode == FRB_GHOST_WAVE_Y) r = rc_ghost_x(p, u, g, p->nx, 1, 1, 1);
  else {
    frb_set_error("slab-parallel ghost fill supports the periodic wave modes");
    return FRB_ERR_ARG;
  }
  return r ? r : 1;
}

// x half of the ghost fill on u plus the frozen ghost columns of rows 1..ny into d1 / d2 (either may be NULL)
int frb_rc_ghost_x_ring(frb_prob_t p, double *u, double *d1, double *d2, int mode) {
  const RcGeom g = geom_of(p);
  if (mode != FRB_GHOST_WAVE_X && mode != FRB_GHOST_WAVE_Y) {
    frb_set_error("slab-parallel ghost fill supports the periodic wave modes");
    return FRB_ERR_ARG;
  }
  dim3 blk(128), grd((g.ny + 2 + 127) / 128, g.nplanes);
  rc_ghost_x_ring_kernel<<<grd, blk, 0, p->ctx->stream>>>(u, d1, d2, g, p->nsp * p->nsp, p->nx, 1,
                                                           mode == FRB_GHOST_WAVE_X ? -1 : 1);
  if (int r = check_launch_rc("rc_ghost_x_ring_kernel")) return r;
  return 1;
}

int frb_rc_row_push3(frb_prob_t p, const double *src, double *const *dst_lo, double *const *dst_hi, int nyl_lo,
                     int flip_var) {
  const RcGeom g = geom_of(p);
  RcPush3 d;
  for (int r = 0; r < 3; ++r) { d.lo[r] = dst_lo ? dst_lo[r] : nullptr; d.hi[r] = dst_hi ? dst_hi[r] : nullptr; }
  dim3 blk(256), grd((unsigned)((g.row + 255) / 256));
  rc_row_push3_kernel<<<grd, blk, 0, p->ctx->stream>>>(src, d, g, nyl_lo, p->nsp * p->nsp, flip_var);
  if (int r = check_launch_rc("rc_row_push3_kernel")) return r;
  return 1;
}

int frb_rc_ring_copy(frb_prob_t p, const double *src, double *dst, bool row0, bool rowN) {
  const RcGeom g = geom_of(p);
  int n = 0;
  if (row0 || rowN) {
    dim3 blk(256), grd((unsigned)((g.row + 255) / 256));
    rc_ring_rows_kernel<<<grd, blk, 0, p->ctx->stream>>>(src, dst, g, row0, rowN);
    if (int r = check_launch_rc("rc_ring_rows_kernel")) return r;
    ++n;
  }
  dim3 blk(128), grd((g.ny + 127) / 128, g.nplanes);
  rc_ring_cols_kernel<<<grd, blk, 0, p->ctx->stream>>>(src, dst, g);
  if (int r = check_launch_rc("rc_ring_cols_kernel")) return r;
  return n + 1;
}

int frb_rc_row_push(frb_prob_t p, const double *src, double *dst_lo, double *dst_hi, int nyl_lo, int flip_var) {
  const RcGeom g = geom_of(p);
  dim3 blk(256), grd((unsigned)((g.row + 255) / 256));
  rc_row_push_kernel<<<grd, blk, 0, p->ctx->stream>>>(src, dst_lo, dst_hi, g, nyl_lo, p->nsp * p->nsp, flip_var);
  if (int r = check_launch_rc("rc_row_push_kernel")) return r;
  return 1;
}
