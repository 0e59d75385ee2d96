// Row-chunk ("RC") device layout of the 2-D Euler state -- the streaming layout of the
// marching stage kernel (frb_euler2d_rc.cu).
//
// The reference's memory image u[i, j, k, l, m] (i fastest, ghosts included) is 4*nsp^2 planes
// of (nx+2)(ny+2) doubles: a CTA that marches over the rows of a 30-element strip touches 64
// separate 240-byte pieces per row and stream, which caps DRAM at ~5.1 TB/s on B200.  In the RC
// layout the same working set is ONE contiguous, 128-byte aligned chunk:
//
//   chunk(j, s) = [plane 0 .. 4*nsp^2-1][lane 0..31]   doubles,   chunks ordered [j][s]
//   lane <-> element column  i = 30*s + lane   (i = 0 and nx+1 are the ghost columns)
//
// so a strip's row moves with one bulk copy (cp.async.bulk, 16 KB at p3) at copy-engine
// bandwidth (measured 6.7-6.9 TB/s for the 16-B / 24-B stage skeletons, scripts/micro/).
// Lanes 1..30 of a chunk are the strip's own elements; lanes 0 and 31 DUPLICATE the edge
// columns of the neighbouring strips (or hold the ghost columns), which is what gives every
// strip its x halo without a second access.  Whoever writes column i writes all its copies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kRcOwn = 30;  // owned element columns per strip

struct RcGeom {
  int nx, ny;       // interior elements
  int ns;           // strips = ceil(nx / 30)
  int nplanes;      // 4 * nsp^2
  int chunk;        // doubles per chunk = nplanes * 32
  size_t row;       // doubles per element row = ns * chunk
  size_t len;       // doubles per buffer = (ny + 2) * row
};

inline RcGeom rc_geom(int nx, int ny, int nsp) {
  RcGeom g;
  g.nx = nx;
  g.ny = ny;
  g.ns = (nx + kRcOwn - 1) / kRcOwn;
  g.nplanes = 4 * nsp * nsp;
  g.chunk = g.nplanes * 32;
  g.row = (size_t)g.ns * g.chunk;
  g.len = (size_t)(ny + 2) * g.row;
  return g;
}

__host__ __device__ inline size_t rc_index(const RcGeom &g, int j, int s, int plane, int lane) {
  return (size_t)j * g.row + (size_t)s * g.chunk + (size_t)plane * 32 + lane;
}

// strip and lane of the primary copy of column i (0 <= i <= nx+1)
__host__ __device__ inline void rc_primary(const RcGeom &g, int i, int *s, int *lane) {
  int s0 = (i - 1) / kRcOwn;  // i = 0 -> 0
  if (s0 < 0) s0 = 0;
  if (s0 > g.ns - 1) s0 = g.ns - 1;
  *s = s0;
  *lane = i - kRcOwn * s0;
}

// the duplicate of column i in a neighbouring strip, if it has one
__host__ __device__ inline bool rc_duplicate(const RcGeom &g, int s, int lane, int *s2, int *lane2) {
  if (lane == 1 && s > 0) { *s2 = s - 1; *lane2 = 31; return true; }
  if (lane == kRcOwn && s < g.ns - 1) { *s2 = s + 1; *lane2 = 0; return true; }
  return false;
}

// Slab-parallel arguments of one launch of the row-chunk stage kernel (frb_euler2d_rc.cu).  The exchange with the
// neighbouring ranks happens INSIDE the stage kernel:
//   * the CTAs that own row 1 / row ny also store it into a slot of the neighbour's halo ring (peer memory over
//     NVLink); when all strips of the row are out, the last CTA raises the neighbour's mailbox (st.release.sys);
//   * the CTAs that need row 0 / row ny+1 poll the local mailbox (ld.acquire.sys) and take the row from the local
//     slot; every other CTA starts at once, so the exchange overlaps the interior rows;
//   * rows 1 and ny are one-row segments scheduled first (blockIdx.y 0 and 1): the flags go up a few
//     microseconds into the launch and the neighbour has a whole stage of slack.
// kRcSlots slots per side: the writer is at most one stage ahead of the reader, who reads the slot of the stage
// before its own -> three live slots.
constexpr int kRcSlots = 4;

struct RcHalo {
  int active;                             // 0: single-domain launch, everything below unused
  const double *src_lo, *src_hi;          // local slot row replacing row 0 / ny+1 of u (nullptr: the array's own row)
  unsigned long long wait_lo, wait_hi;    // mailbox value to wait for before reading src_lo / src_hi
  unsigned long long *mailbox;            // local mailbox: [0] from lo, [1] from hi, [2] time-out record
  double *dst_lo, *dst_hi;                // peer slot row for my row 1 / ny (nullptr: global seam, no data)
  unsigned long long *flag_lo, *flag_hi;  // the neighbours' mailbox words to raise
  unsigned long long epoch;               // value to raise them to
  unsigned int *count;                    // local completion counters [2] (monotonic, +ns per launch)
};
