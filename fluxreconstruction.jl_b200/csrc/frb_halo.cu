// Multi-GPU halo exchange for the row-slab partition of the 2-D problems.
//
// One process per GPU.  Each rank owns ny_local rows; local rows 0 and ny_local+1 are halo
// rows.  Between two ranks the halo row holds the neighbour's boundary row of the *current*
// stage; across the global y seam it is the reference's frozen per-step ghost row
// (example/euler2d_wave.jl:127-132).  Rows travel by direct stores into the neighbour's
// memory over NVLink (the buffers are mapped with CUDA IPC): the marching stage kernel
// writes its first/last owned row to the peer while it writes it locally, a one-thread
// kernel then raises a flag in the peer's mailbox (st.release.sys), and the peer's next
// stage is preceded by a one-thread wait kernel (ld.acquire.sys).  No host round trip, no
// collective: the path's only exchange is nearest-neighbour.
#include <cstdlib>
#include <cstring>

#include "frb_internal.cuh"
#include "frb_rc.cuh"

struct FrbHalo {
  int rank = 0, nranks = 1;
  int nyl_lo = 0, nyl_hi = 0;          // owned rows of the neighbours
  double *peer_lo[3] = {nullptr, nullptr, nullptr};  // neighbour below: its u, s1, s2
  double *peer_hi[3] = {nullptr, nullptr, nullptr};
  double *rc_lo[3] = {nullptr, nullptr, nullptr};    // the same three roles in the row-chunk layout
  double *rc_hi[3] = {nullptr, nullptr, nullptr};
  unsigned long long *flags = nullptr;     // local mailbox: [0] from lo, [1] from hi, [2] error
  unsigned long long *flags_lo = nullptr;  // neighbours' mailboxes
  unsigned long long *flags_hi = nullptr;
  unsigned long long epoch = 0;
  // in-kernel exchange of the row-chunk stage kernel (RcHalo, frb_rc.cuh): a ring of kRcSlots halo rows per side,
  // written by the neighbours' stage kernels; slots = [side lo | hi][kRcSlots][row]
  double *slots = nullptr, *slots_lo = nullptr, *slots_hi = nullptr;  // local ring, the neighbours' rings
  unsigned int *counters = nullptr;     // [2] strips that have finished row 1 / row ny (monotonic)
  long long seq = 0;                    // in-kernel stages launched so far (ring position)
  long long slot_seq[3] = {-1, -1, -1}; // per buffer role: the stage whose ring slot holds its halo rows, -1 = the
  unsigned long long slot_epoch[3] = {0, 0, 0};  //   array's own rows 0 / ny+1 do; and that stage's mailbox epoch
  void *opened[12] = {nullptr};
  int nopened = 0;
};

namespace {

// copy full rows (all NXG columns, every plane) into the neighbours' halo rows
__global__ void halo_push_kernel(const double *__restrict__ src, double *__restrict__ dst_lo,
                                 double *__restrict__ dst_hi, int nx, int nyl, int nyl_lo, int nyl_hi,
                                 int nplanes, int npp, int flip_var_lo, int flip_var_hi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int p = blockIdx.y;
  if (i > nx + 1 || p >= nplanes) return;
  const size_t NXG = nx + 2;
  const size_t NE = NXG * (size_t)(nyl + 2);
  const int var = p / npp;
  if (dst_lo) {  // my first owned row -> upper halo row of the rank below
    const size_t NEl = NXG * (size_t)(nyl_lo + 2);
    double v = src[i + NXG * 1 + NE * p];
    dst_lo[i + NXG * (size_t)(nyl_lo + 1) + NEl * p] = (var == flip_var_lo) ? -v : v;
  }
  if (dst_hi) {  // my last owned row -> lower halo row of the rank above
    const size_t NEh = NXG * (size_t)(nyl_hi + 2);
    double v = src[i + NXG * (size_t)nyl + NE * p];
    dst_hi[i + NEh * p] = (var == flip_var_hi) ? -v : v;
  }
}

// cfg5 (u[4, nsp, nsp, ny+2, nx+2], slabs along the slowest index i): a column of cells is one contiguous run
// of `col` doubles.  My first owned column -> the upper halo column of the rank below, my last owned column -> the
// lower halo column of the rank above.
__global__ void halo_push_cols_kernel(const double *__restrict__ src, double *__restrict__ dst_lo,
                                      double *__restrict__ dst_hi, size_t col, int nxl, int nxl_lo) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= col) return;
  if (dst_lo) dst_lo[col * (size_t)(nxl_lo + 1) + t] = src[col * 1 + t];
  if (dst_hi) dst_hi[t] = src[col * (size_t)nxl + t];
}

__global__ void halo_signal_kernel(unsigned long long *flag_lo, unsigned long long *flag_hi,
                                   unsigned long long value) {
  __threadfence_system();
  if (flag_lo) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_lo), "l"(value) : "memory");
  if (flag_hi) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag_hi), "l"(value) : "memory");
}

__global__ void halo_wait_kernel(unsigned long long *flags, int wait_lo, int wait_hi,
                                 unsigned long long value, unsigned long long timeout_ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int s = 0; s < 2; ++s) {
    if (!(s == 0 ? wait_lo : wait_hi)) continue;
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + s) : "memory");
      if (v >= value) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {  // never hang the box: record and carry on
        flags[2] = value;
        break;
      }
      __nanosleep(200);
    } while (true);
  }
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, what, __FILE__, __LINE__);
  return 0;
}

}  // namespace

// export layout: 6 IPC handles (u, s1, s2, flags, row-chunk buffers, halo ring) + int32 ny_local + int32 has_rc
extern "C" int32_t frb_halo_export(frb_prob_t p, unsigned char *out) {
  if (!p || !out) { frb_set_error("frb_halo_export: NULL argument"); return FRB_ERR_ARG; }
  // curvilinear euler2d problems take the same row slabs (their metric, normals and factor tables are per-rank
  // data of frb_euler2d_curv_create; the stage kernels read rows 0 / ny+1 of the state like any other row)
  if (!(p->kind == K_EULER2D || p->kind == K_NS2D)) {
    frb_set_error("frb_halo_export: euler2d and ns2d problems only");
    return FRB_ERR_STATE;
  }
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  if (!p->halo) {
    p->halo = new FrbHalo();
    FRB_CUDA(cudaMalloc(&p->halo->flags, 4 * sizeof(unsigned long long)));
    FRB_CUDA(cudaMemset(p->halo->flags, 0, 4 * sizeof(unsigned long long)));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == FRB_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  void *bufs[4] = {p->u, p->s1, p->s2, p->halo->flags};
  for (int b = 0; b < 4; ++b) {
    FRB_CUDA(cudaIpcGetMemHandle(&h, bufs[b]));
    memcpy(out + b * FRB_IPC_HANDLE_BYTES, &h, FRB_IPC_HANDLE_BYTES);
  }
  // extent of the slab along the partitioned index: rows (euler2d, slabs along j) or columns (ns2d, along i)
  int32_t tail[2] = {p->kind == K_NS2D ? p->nx : p->ny, p->rc_base ? 1 : 0};
  memset(out + 4 * FRB_IPC_HANDLE_BYTES, 0, FRB_IPC_HANDLE_BYTES);
  if (p->rc_base) {
    FRB_CUDA(cudaIpcGetMemHandle(&h, p->rc_base));
    memcpy(out + 4 * FRB_IPC_HANDLE_BYTES, &h, FRB_IPC_HANDLE_BYTES);
  }
  memset(out + 5 * FRB_IPC_HANDLE_BYTES, 0, FRB_IPC_HANDLE_BYTES);
  if (p->rc_base) {
    FrbHalo *H = p->halo;
    if (!H->slots) {
      const size_t n = (size_t)2 * kRcSlots * rc_geom(p->nx, p->ny, p->nsp).row;
      FRB_CUDA(cudaMalloc(&H->slots, sizeof(double) * n));
      FRB_CUDA(cudaMemset(H->slots, 0, sizeof(double) * n));
      FRB_CUDA(cudaMalloc(&H->counters, 2 * sizeof(unsigned int)));
      FRB_CUDA(cudaMemset(H->counters, 0, 2 * sizeof(unsigned int)));
    }
    FRB_CUDA(cudaIpcGetMemHandle(&h, H->slots));
    memcpy(out + 5 * FRB_IPC_HANDLE_BYTES, &h, FRB_IPC_HANDLE_BYTES);
  }
  memcpy(out + 6 * FRB_IPC_HANDLE_BYTES, tail, sizeof tail);
  return FRB_OK;
}

static int open_peer(frb_prob_t p, const unsigned char *blob, double **bufs3, double **rc3,
                     unsigned long long **flags, int *nyl, double **ring) {
  FrbHalo *H = p->halo;
  for (int b = 0; b < 4; ++b) {
    cudaIpcMemHandle_t h;
    memcpy(&h, blob + b * FRB_IPC_HANDLE_BYTES, FRB_IPC_HANDLE_BYTES);
    void *ptr = nullptr;
    FRB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    H->opened[H->nopened++] = ptr;
    if (b < 3) bufs3[b] = static_cast<double *>(ptr);
    else *flags = static_cast<unsigned long long *>(ptr);
  }
  int32_t tail[2];
  memcpy(tail, blob + 6 * FRB_IPC_HANDLE_BYTES, sizeof tail);
  *nyl = tail[0];
  *ring = nullptr;
  if (tail[1] && p->rc_base) {  // the neighbour's ru | rs1 | rs2, sized for ITS row count
    cudaIpcMemHandle_t h;
    memcpy(&h, blob + 4 * FRB_IPC_HANDLE_BYTES, FRB_IPC_HANDLE_BYTES);
    void *ptr = nullptr;
    FRB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    H->opened[H->nopened++] = ptr;
    const size_t len = rc_geom(p->nx, tail[0], p->nsp).len;
    for (int b = 0; b < 3; ++b) rc3[b] = static_cast<double *>(ptr) + b * len;
    memcpy(&h, blob + 5 * FRB_IPC_HANDLE_BYTES, FRB_IPC_HANDLE_BYTES);
    ptr = nullptr;
    FRB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    H->opened[H->nopened++] = ptr;
    *ring = static_cast<double *>(ptr);
  } else if (p->rc_base) {
    frb_set_error("frb_halo_connect: neighbour has no row-chunk buffers (mixed configurations)");
    return FRB_ERR_PEER;
  }
  return FRB_OK;
}

extern "C" int32_t frb_halo_connect(frb_prob_t p, int32_t rank, int32_t nranks,
                                    const unsigned char *blob_lo, const unsigned char *blob_hi) {
  if (!p || !p->halo) { frb_set_error("frb_halo_connect: call frb_halo_export first"); return FRB_ERR_STATE; }
  if (nranks < 2 || !blob_lo || !blob_hi) {
    frb_set_error("frb_halo_connect: needs >= 2 ranks and both neighbour blobs");
    return FRB_ERR_ARG;
  }
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  FrbHalo *H = p->halo;
  H->rank = rank;
  H->nranks = nranks;
  if (int rc = open_peer(p, blob_lo, H->peer_lo, H->rc_lo, &H->flags_lo, &H->nyl_lo, &H->slots_lo)) return rc;
  if (nranks == 2) {  // both neighbours are the same process: map once
    for (int b = 0; b < 3; ++b) { H->peer_hi[b] = H->peer_lo[b]; H->rc_hi[b] = H->rc_lo[b]; }
    H->flags_hi = H->flags_lo;
    H->nyl_hi = H->nyl_lo;
    H->slots_hi = H->slots_lo;
  } else {
    if (int rc = open_peer(p, blob_hi, H->peer_hi, H->rc_hi, &H->flags_hi, &H->nyl_hi, &H->slots_hi)) return rc;
  }
  return frb_halo_sync(p);
}

// (re)send this rank's boundary rows of the resident state to the interior neighbours; every
// rank must call it after (re)uploading its slab (collective in effect, not in mechanism)
extern "C" int32_t frb_halo_sync(frb_prob_t p) {
  if (!p || !frb_halo_active(p)) { frb_set_error("frb_halo_sync: halo not connected"); return FRB_ERR_STATE; }
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  int n = frb_halo_push(p, p->u, 0, false, -1);
  if (n < 0) return n;
  if ((n = frb_halo_signal(p)) < 0) return n;
  p->halo_pending = true;
  return FRB_OK;
}

extern "C" int32_t frb_halo_disconnect(frb_prob_t p) {
  if (!p || !p->halo) return FRB_OK;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  FrbHalo *H = p->halo;
  for (int q = 0; q < H->nopened; ++q) cudaIpcCloseMemHandle(H->opened[q]);
  cudaFree(H->flags);
  cudaFree(H->slots);
  cudaFree(H->counters);
  delete H;
  p->halo = nullptr;
  return FRB_OK;
}

int frb_halo_rank(frb_prob_t p, int *nranks) {
  if (nranks) *nranks = p->halo ? p->halo->nranks : 1;
  return p->halo ? p->halo->rank : 0;
}

bool frb_halo_active(frb_prob_t p) { return p->halo && p->halo->nranks > 1 && p->halo->flags_lo; }

// which local buffer is `ptr`?  (roles move with the pointer swaps of the Euler scheme)
static bool is_rc(frb_prob_t p, const double *ptr) {
  return p->rc_base && (ptr == p->ru || ptr == p->rs1 || ptr == p->rs2);
}
static int role_of(frb_prob_t p, const double *ptr) {
  if (is_rc(p, ptr)) return ptr == p->ru ? 0 : ptr == p->rs1 ? 1 : 2;
  return ptr == p->u ? 0 : ptr == p->s1 ? 1 : ptr == p->s2 ? 2 : -1;
}

void frb_halo_swap_roles(frb_prob_t p, int a, int b, bool rc) {
  if (!p->halo) return;
  if (rc) {
    std::swap(p->halo->rc_lo[a], p->halo->rc_lo[b]);
    std::swap(p->halo->rc_hi[a], p->halo->rc_hi[b]);
    std::swap(p->halo->slot_seq[a], p->halo->slot_seq[b]);
    std::swap(p->halo->slot_epoch[a], p->halo->slot_epoch[b]);
  } else {
    std::swap(p->halo->peer_lo[a], p->halo->peer_lo[b]);
    std::swap(p->halo->peer_hi[a], p->halo->peer_hi[b]);
  }
}

// peer destinations of the interior slab boundaries for a stage writing `out`
void frb_halo_stage_targets(frb_prob_t p, const double *out, double **dst_lo, double **dst_hi,
                            int *nyl_lo, int *nyl_hi) {
  *dst_lo = *dst_hi = nullptr;
  *nyl_lo = *nyl_hi = 0;
  if (!frb_halo_active(p)) return;
  FrbHalo *H = p->halo;
  int r = role_of(p, out);
  if (r < 0) return;
  const bool rc = is_rc(p, out);
  if (H->rank != 0) { *dst_lo = rc ? H->rc_lo[r] : H->peer_lo[r]; *nyl_lo = H->nyl_lo; }
  if (H->rank != H->nranks - 1) { *dst_hi = rc ? H->rc_hi[r] : H->peer_hi[r]; *nyl_hi = H->nyl_hi; }
}

int frb_halo_role(frb_prob_t p, const double *ptr) { return role_of(p, ptr); }

// Push the boundary rows of `src` into role `dst_role` (0 = u, 1 = s1, 2 = s2) of the
// neighbours.  seam = false: interior slab boundaries only (per stage); seam = true: only the
// global y seam (per step, frozen ghost rows) with the sign flip of the ghost mode.
int frb_halo_push(frb_prob_t p, const double *src, int dst_role, bool seam, int flip_var) {
  if (!frb_halo_active(p)) return 0;
  FrbHalo *H = p->halo;
  if (dst_role < 0 || dst_role > 2) { frb_set_error("halo push: unknown buffer"); return FRB_ERR_STATE; }
  const bool first = H->rank == 0, last = H->rank == H->nranks - 1;
  const bool rc = is_rc(p, src);
  double **plo = rc ? H->rc_lo : H->peer_lo, **phi = rc ? H->rc_hi : H->peer_hi;
  double *dl = nullptr, *dh = nullptr;
  if (!seam) {
    if (!first) dl = plo[dst_role];
    if (!last) dh = phi[dst_role];
  } else {
    if (first) dl = plo[dst_role];
    if (last) dh = phi[dst_role];
  }
  if (!dl && !dh) return 0;
  if (p->kind == K_NS2D) {  // no periodic seam in the cavity: interior slab boundaries only
    if (seam) return 0;
    const size_t col = (size_t)4 * p->nsp * p->nsp * (p->ny + 2);
    halo_push_cols_kernel<<<(unsigned)((col + 255) / 256), 256, 0, p->ctx->stream>>>(src, dl, dh, col, p->nx,
                                                                                     H->nyl_lo);
    if (int r = check_launch("halo_push_cols_kernel")) return r;
    return 1;
  }
  if (rc) {
    if (!seam) H->slot_seq[dst_role] = -1;  // every rank pushes: the interior halo rows of this role are in the arrays
    return frb_rc_row_push(p, src, dl, dh, H->nyl_lo, seam ? flip_var : -1);
  }
  const int npp = p->nsp * p->nsp, nplanes = 4 * npp;
  dim3 blk(128), grd((p->nx + 2 + 127) / 128, nplanes);
  halo_push_kernel<<<grd, blk, 0, p->ctx->stream>>>(src, dl, dh, p->nx, p->ny, H->nyl_lo, H->nyl_hi,
                                                    nplanes, npp, seam ? flip_var : -1,
                                                    seam ? flip_var : -1);
  if (int rc = check_launch("halo_push_kernel")) return rc;
  return 1;
}

// the frozen seam rows of a step into all three buffers of the seam neighbour: one launch on the row-chunk path
int frb_halo_push_seam_all(frb_prob_t p, const double *src, int flip_var) {
  if (!frb_halo_active(p)) return 0;
  FrbHalo *H = p->halo;
  const bool first = H->rank == 0, last = H->rank == H->nranks - 1;
  if (!first && !last) return 0;
  if (!is_rc(p, src)) {
    int n = 0;
    for (int role = 0; role < 3; ++role) {
      int r = frb_halo_push(p, src, role, true, flip_var);
      if (r < 0) return r;
      n += r;
    }
    return n;
  }
  return frb_rc_row_push3(p, src, first ? H->rc_lo : nullptr, last ? H->rc_hi : nullptr, H->nyl_lo, flip_var);
}

// raise this rank's epoch in both neighbours' mailboxes, after everything queued so far
int frb_halo_signal(frb_prob_t p) {
  if (!frb_halo_active(p)) return 0;
  FrbHalo *H = p->halo;
  H->epoch += 1;
  // I am the "hi" neighbour of the rank below and the "lo" neighbour of the rank above
  halo_signal_kernel<<<1, 1, 0, p->ctx->stream>>>(H->flags_lo + 1, H->flags_hi + 0, H->epoch);
  if (int rc = check_launch("halo_signal_kernel")) return rc;
  p->halo_pending_legacy = true;
  return 1;
}

// ---- in-kernel exchange of the row-chunk stage kernel ---------------------------------------------------
bool frb_halo_rc_inkernel(frb_prob_t p) {
  static const bool off = getenv("FRB_HALO_LEGACY") != nullptr;  // measurement switch: the round-1 epoch kernels
  return !off && frb_halo_active(p) && p->halo->slots && p->halo->slots_lo && p->halo->slots_hi;
}

void frb_halo_rc_reset(frb_prob_t p) {
  if (!p->halo) return;
  for (int r = 0; r < 3; ++r) p->halo->slot_seq[r] = -1;
}

// do the interior halo rows of RC buffer u live in the array itself (after an upload / conversion / push)?
bool frb_halo_rc_input_in_array(frb_prob_t p, const double *u) {
  const int r = role_of(p, u);
  return r < 0 || p->halo->slot_seq[r] < 0;
}

// Arguments of one stage launch u -> out: where the input's halo rows are, which ring slot of the neighbours takes
// this stage's boundary rows, the mailbox epoch that announces them.  Advances the epoch and the ring position
// (every rank launches the same sequence of stages, so the counters agree without any communication).
int frb_halo_rc_stage(frb_prob_t p, const double *u, const double *out, RcHalo *h) {
  memset(h, 0, sizeof *h);
  if (!frb_halo_rc_inkernel(p)) return 0;
  FrbHalo *H = p->halo;
  const int ru = role_of(p, u), ro = role_of(p, out);
  if (ru < 0 || ro < 0 || !is_rc(p, u) || !is_rc(p, out)) {
    frb_set_error("halo: stage buffers are not the problem's row-chunk buffers");
    return FRB_ERR_STATE;
  }
  const size_t row = rc_geom(p->nx, p->ny, p->nsp).row;
  const bool first = H->rank == 0, last = H->rank == H->nranks - 1;
  h->active = 1;
  h->mailbox = H->flags;
  if (H->slot_seq[ru] >= 0) {
    const size_t s = (size_t)(H->slot_seq[ru] % kRcSlots);
    if (!first) { h->src_lo = H->slots + (0 * kRcSlots + s) * row; h->wait_lo = H->slot_epoch[ru]; }
    if (!last) { h->src_hi = H->slots + (1 * kRcSlots + s) * row; h->wait_hi = H->slot_epoch[ru]; }
  }
  H->epoch += 1;
  H->seq += 1;
  const size_t s = (size_t)(H->seq % kRcSlots);
  if (!first) h->dst_lo = H->slots_lo + (1 * kRcSlots + s) * row;  // my row 1 -> the hi side of the rank below
  if (!last) h->dst_hi = H->slots_hi + (0 * kRcSlots + s) * row;   // my row ny -> the lo side of the rank above
  h->flag_lo = H->flags_lo + 1;
  h->flag_hi = H->flags_hi + 0;
  h->epoch = H->epoch;
  h->count = H->counters;
  H->slot_seq[ro] = H->seq;
  H->slot_epoch[ro] = H->epoch;
  return 1;
}

// bring the interior halo rows of U from the ring into the array (rows 0 / ny+1), for everything that reads the
// array as a whole (download, conversion to the reference image, f! of the resident slab).  The caller has waited
// for the epoch.
int frb_halo_rc_flush(frb_prob_t p, double *U) {
  if (!frb_halo_active(p) || !p->halo->slots) return 0;
  FrbHalo *H = p->halo;
  const int r = role_of(p, U);
  if (r < 0 || H->slot_seq[r] < 0) return 0;
  const RcGeom g = rc_geom(p->nx, p->ny, p->nsp);
  const size_t s = (size_t)(H->slot_seq[r] % kRcSlots);
  int n = 0;
  if (H->rank != 0) {
    FRB_CUDA(cudaMemcpyAsync(U, H->slots + (0 * kRcSlots + s) * g.row, sizeof(double) * g.row,
                             cudaMemcpyDeviceToDevice, p->ctx->stream));
    ++n;
  }
  if (H->rank != H->nranks - 1) {
    FRB_CUDA(cudaMemcpyAsync(U + g.row * (size_t)(g.ny + 1), H->slots + (1 * kRcSlots + s) * g.row,
                             sizeof(double) * g.row, cudaMemcpyDeviceToDevice, p->ctx->stream));
    ++n;
  }
  H->slot_seq[r] = -1;
  return n;
}

// block the stream until both neighbours have reached this rank's current epoch
int frb_halo_wait(frb_prob_t p) {
  if (!frb_halo_active(p)) return 0;
  FrbHalo *H = p->halo;
  halo_wait_kernel<<<1, 1, 0, p->ctx->stream>>>(H->flags, 1, 1, H->epoch, 5000000000ull);
  if (int rc = check_launch("halo_wait_kernel")) return rc;
  return 1;
}

int frb_halo_check_timeout(frb_prob_t p) {
  if (!frb_halo_active(p)) return 0;
  unsigned long long err = 0;
  FRB_CUDA(cudaMemcpy(&err, p->halo->flags + 2, sizeof err, cudaMemcpyDeviceToHost));
  if (err) {
    frb_set_error("halo wait timed out at epoch " + std::to_string(err) + " (a neighbour rank stalled)");
    return FRB_ERR_PEER;
  }
  return 0;
}
