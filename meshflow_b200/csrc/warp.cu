// Mesh warp + crop on the device (hot-path subsystem 3).
//
// The reference loops over the R*C mesh cells of every frame and, for each cell, runs ~10 full-frame
// NumPy/OpenCV passes (mfs.py:1031-1061), then one cv2.remap (mfs.py:1063).  Here:
//
//   cell_setup_kernel : one thread per (frame, cell): the two exact 4-point homographies, the
//                       adjugate inverse used by the membership test, the integer bounds of the rest
//                       cell and the cell's support box in the output frame; the thread then files
//                       the cell into the candidate list of every 64x16 output tile its box touches.
//   warp_kernel       : one CTA per 128x8 output tile, one warp per row segment, four adjacent pixels
//                       per thread.  Sorts the tile's candidates by descending cell id ("the last
//                       cell written wins", mfs.py:1060-1061); per pixel the first candidate whose
//                       warped rectangle mask is non-zero wins.  Membership is screened in float32
//                       (cell-local coordinates, conservative margin) and only decided in float64 --
//                       with the reference's exact operation order -- inside the margin; the remap
//                       coordinate is always the reference's float64 sequence.  The four taps use
//                       OpenCV's 1/32-px fixed-point weights; when the four pixels' footprints are
//                       adjacent the two source rows are read as aligned 32-bit words.  The four
//                       crop-edge searches (mfs.py:1075-1098) are warp reductions + atomics.
//   crop_resize_kernel: cv2.resize of the cropped window back to W x H (mfs.py:1150-1155).
//
// warp_kernel is the GENERIC pixel kernel (float32 maps for parity tests, tiny frames); the production path --
// row segments, analytic crop edges, the fast and the fused pixel kernels -- is in warp_fast.cuh.
//
// HBM layout: frames [nf][H][W][3] uint8 (row pitch 3W, no padding -- BGR24 rows are multiples of 16
// bytes at every standard resolution); cells [nf][R*C] mf::Cell (240 B); tile lists
// [nf][tiles][kTileCap] uint16 + counts.
#include <stdlib.h>
#include "mf_common.cuh"
#include "mf_math.cuh"

namespace mf {

static constexpr int kTileW = 128;       // warp kernel: one warp = 128 x 1 output pixels, 4 per thread
static constexpr int kTileH = 8;
static constexpr int kPix = 4;
static constexpr int kTileCap = 48;      // candidate cells per tile before the exhaustive fallback
static constexpr int kCountMask = 0xffff; // tile_count: low 16 bits = candidates, bit 16 = has a border cell
static constexpr int kEdgeFlag = 0x10000;
static constexpr int kWarpThreads = 256; // 32 lanes x 8 rows
#ifndef MF_FAST_ROWS
#define MF_FAST_ROWS 15
#endif
static constexpr int kFastRows = MF_FAST_ROWS;     // fast path: rows per warp (<= 16: one 64-bit mask of 4 bits per row)
static constexpr int kFastTileH = 8 * kFastRows;   // 120: divides 720, 1080, 1440, 2160, 4320
static_assert(kFastRows >= 1 && kFastRows <= 16, "four mask bits per row in one 64-bit word, four row bits per queue entry");

// Exact 4-point homography (mf_math.cuh homography_4pt: 8x8 Gaussian elimination with partial pivoting, fixed
// operation order) with EIGHT lanes per system, one matrix row each: pivot search and row exchange by shuffles, every
// lane eliminates its own row.  Same operations on the same operands as the sequential routine, so the result is
// bit-identical; the sequential version kept the 8x9 matrix in local memory behind a data-dependent pivot index and
// ran at 6 % occupancy.  Returns h[0..7] (h22 = 1) in every lane of the group.
__device__ __forceinline__ double shfl8(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }

// (x, y) -> (X, Y): the correspondence of THIS lane's corner (corner = r >> 1).
__device__ __forceinline__ void homography_4pt_rows(double x, double y, double X, double Y, int r, double (&h)[8]) {
  double a[9];
  {
    const double D = (r & 1) ? Y : X;
    const bool odd = (r & 1) != 0;
    a[0] = odd ? 0.0 : x; a[1] = odd ? 0.0 : y; a[2] = odd ? 0.0 : 1.0;
    a[3] = odd ? x : 0.0; a[4] = odd ? y : 0.0; a[5] = odd ? 1.0 : 0.0;
    a[6] = -MF_MUL(x, D); a[7] = -MF_MUL(y, D); a[8] = D;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // partial pivoting: the FIRST row >= k with the largest |a[k]| (strict comparison in the sequential routine)
    double best = r >= k ? fabs(a[k]) : -1.0;
    int piv = r;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, m, 8);
      const int op = __shfl_xor_sync(0xffffffffu, piv, m, 8);
      if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
    }
    double p[9];
#pragma unroll
    for (int j = k; j < 9; ++j) {
      const double from_piv = shfl8(a[j], piv), from_k = shfl8(a[j], k);
      p[j] = from_piv;                                       // the pivot row, known to every lane
      if (r == k) a[j] = from_piv; else if (r == piv) a[j] = from_k;
    }
    if (r > k) {
      const double f = MF_DIV(a[k], p[k]);
#pragma unroll
      for (int j = k; j < 9; ++j) a[j] = MF_SUB(a[j], MF_MUL(f, p[j]));
    }
  }
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    double acc = a[8];
#pragma unroll
    for (int j = i + 1; j < 8; ++j) acc = MF_SUB(acc, MF_MUL(a[j], h[j]));
    h[i] = shfl8(MF_DIV(acc, a[i]), i);                      // row i's lane holds the valid value
  }
}

// Sixteen lanes per cell: lanes 0-7 solve the unstabilized -> stabilized homography, lanes 8-15 the inverse
// direction (a separate solve, mfs.py:1041-1042).  homs[cell] = Hus[0..7], Hsu[0..7].
__global__ void __launch_bounds__(128) cell_homographies_kernel(
    const double* __restrict__ u, const double* __restrict__ s, const float* __restrict__ vertex_xy,
    int nf, int R, int C, double* __restrict__ homs) {
  const int ncell = R * C;
  const int64_t total = (int64_t)nf * ncell;
  const int sub = threadIdx.x & 15;
  int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const bool valid = idx < total;
  if (!valid) idx = total - 1;                               // keep the shuffles of the warp complete
  const int f = (int)(idx / ncell), id = (int)(idx - (int64_t)f * ncell);
  const int r = id / C, c = id - r * C;
  const int V = (R + 1) * (C + 1);
  // this lane's matrix row belongs to corner (row & 7) >> 1 of the cell (TL, TR, BL, BR): only that vertex is read
  const int row = sub & 7, corner = row >> 1;
  const int v = (r + (corner >> 1)) * (C + 1) + c + (corner & 1);
  const size_t o = ((size_t)f * V + v) * 2;
  const double rx = (double)vertex_xy[2 * v], ry = (double)vertex_xy[2 * v + 1];
  // stabilized vertex = rest + (s - u), float64, then rounded to float32 by cv2.findHomography (mfs.py:964-967, 1025)
  const double sx = (double)(float)MF_ADD(rx, MF_SUB(s[o], u[o])), sy = (double)(float)MF_ADD(ry, MF_SUB(s[o + 1], u[o + 1]));
  const bool inverse = sub >= 8;
  double h[8];
  homography_4pt_rows(inverse ? sx : rx, inverse ? sy : ry, inverse ? rx : sx, inverse ? ry : sy, row, h);
  double mine = h[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mine = (sub & 7) == i ? h[i] : mine;
  if (valid) homs[idx * 16 + sub] = mine;
}

// One thread per (frame, cell): the rest of the set-up from the two homographies (adjugate inverse, bounds, support
// box, float32 forms, membership half-planes); the thread then files the cell into the candidate list of every
// 128 x 8 output tile its box touches.
__global__ void __launch_bounds__(128) cell_setup_kernel(
    const double* __restrict__ u, const double* __restrict__ s, const float* __restrict__ vertex_xy,
    const double* __restrict__ homs, int nf, int W, int H, int R, int C, int tiles_x, int tiles_y, Cell* __restrict__ cells,
    CellFast* __restrict__ fast, CellSpan* __restrict__ spans,
    int* __restrict__ tile_count, uint16_t* __restrict__ tile_list, int32_t* __restrict__ crop_out) {
  const int ncell = R * C;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nf * ncell) return;
  const int f = (int)(idx / ncell), id = (int)(idx - (int64_t)f * ncell);
  const int r = id / C, c = id - r * C;
  if (id == 0 && crop_out) {  // identities of the max/max/min/min edge searches (mfs.py:992-995)
    crop_out[4 * f + 0] = 0; crop_out[4 * f + 1] = 0;
    crop_out[4 * f + 2] = W - 1; crop_out[4 * f + 3] = H - 1;
  }
  double rest[8];
  const int vidx[4] = {r * (C + 1) + c, r * (C + 1) + c + 1, (r + 1) * (C + 1) + c, (r + 1) * (C + 1) + c + 1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    rest[2 * k] = (double)vertex_xy[2 * vidx[k]];
    rest[2 * k + 1] = (double)vertex_xy[2 * vidx[k] + 1];
  }
  double Hus[9], Hsu[9];
  const double2* hp = reinterpret_cast<const double2*>(homs + idx * 16);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 a = __ldg(hp + i), b = __ldg(hp + 4 + i);
    Hus[2 * i] = a.x; Hus[2 * i + 1] = a.y; Hsu[2 * i] = b.x; Hsu[2 * i + 1] = b.y;
  }
  Hus[8] = 1.0; Hsu[8] = 1.0;
  Cell cell;
  cell_setup_from_homographies(rest, Hus, Hsu, W, H, cell);
  if (fast != nullptr) {                  // fast path: box-local float32 map + membership half-planes
    CellFast cf;
    CellSpan sp;
    // rest corners are TL, TR, BL, BR of an axis-aligned rectangle (mfs.py:1039-1044)
    cell_fast_setup(cell, (int)floor(fmin(rest[0], rest[4])), (int)ceil(fmax(rest[2], rest[6])),
                    (int)floor(fmin(rest[1], rest[3])), (int)ceil(fmax(rest[5], rest[7])), W, H, cf, sp);
    fast[idx] = cf;
    spans[idx] = sp;
    if (cf.thr_u >= 0.0f) cell.edge_flags |= kMapMonotone;   // the remap denominator keeps its sign over the box
  }
  cells[idx] = cell;
  if (cell.bx0 > cell.bx1) return;
  // Tiles that hold a border cell are tagged: only their rows take part in the crop-edge searches.
  const bool edge_cell = (cell.edge_flags & kEdgeAny) != 0u;
  const int ntiles = tiles_x * tiles_y;
  const int tx0 = cell.bx0 / kTileW, ty0 = cell.by0 / kTileH;
  const int ntx = cell.bx1 / kTileW - tx0 + 1, n = ntx * (cell.by1 / kTileH - ty0 + 1);
  const size_t tbase = (size_t)f * ntiles;
  const int flag = edge_cell ? kEdgeFlag : 0;               // carried by the same atomic as the count
  // four tiles per step: the atomics of a step are independent, so their round trips to L2 overlap
  for (int k0 = 0; k0 < n; k0 += 4) {
    size_t t[4];
    int slot[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = min(k0 + j, n - 1);
      t[j] = tbase + (size_t)(ty0 + k / ntx) * tiles_x + (tx0 + k % ntx);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) slot[j] = k0 + j < n ? (atomicAdd(&tile_count[t[j]], 1) & kCountMask) : kTileCap;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k0 + j >= n) continue;
      if (slot[j] < kTileCap) tile_list[t[j] * kTileCap + slot[j]] = (uint16_t)id;
      if (flag) atomicOr(&tile_count[t[j]], flag);          // no return value needed: tile_sort_kernel lists the tagged tiles
    }
  }
}

// Candidate lists come out of cell_setup_kernel in atomic order; the warp kernel needs them by
// descending cell id ("the last cell written wins", mfs.py:1060-1061).  Lists are short (typically
// 4-8 ids): one thread sorts one list in place.
__global__ void __launch_bounds__(128) tile_sort_kernel(const int* __restrict__ tile_count,
                                                        uint16_t* __restrict__ tile_list, int64_t ntiles_total, int ntiles,
                                                        int* __restrict__ edge_count, uint16_t* __restrict__ edge_tiles) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = t < ntiles_total;
  const int craw = in_range ? tile_count[t] : 0;
  // tiles that hold a border cell: crop_edges_kernel's work list.  One atomic per frame and warp: the lanes of a
  // warp that tag tiles of the same frame share it.
  const int f = in_range ? (int)(t / ntiles) : -1;
  const bool tagged = in_range && (craw & kEdgeFlag) && edge_count != nullptr;
  const unsigned peers = __match_any_sync(0xffffffffu, tagged ? f : -1 - (int)(threadIdx.x & 31));
  if (tagged) {
    const int lane = threadIdx.x & 31, leader = __ffs((int)peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&edge_count[f], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    edge_tiles[(size_t)f * ntiles + base + __popc(peers & ((1u << lane) - 1u))] = (uint16_t)(t - (int64_t)f * ntiles);
  }
  if (!in_range) return;
  const int n = craw & kCountMask;
  if (n < 2 || n > kTileCap) return;
  uint16_t* l = tile_list + t * kTileCap;
  for (int i = 1; i < n; ++i) {
    const uint16_t v = l[i];
    int j = i - 1;
    while (j >= 0 && l[j] < v) { l[j + 1] = l[j]; --j; }
    l[j + 1] = v;
  }
}

// Per-pixel (general) resolution of one output pixel against a list of candidate cells (descending
// id; list == nullptr: every cell of the frame).  Rare path: group straddles a cell edge, cell not
// screenable, frame border.
__device__ __forceinline__ float2 resolve_pixel(const Cell* __restrict__ fcells, const uint16_t* __restrict__ list,
                                             int n, int px, int py, float mx, float my) {
  const double x = (double)px, y = (double)py;
  for (int k = 0; k < n; ++k) {
    const Cell& c = fcells[list ? (int)__ldg(list + k) : n - 1 - k];
    const int4 box = __ldg(reinterpret_cast<const int4*>(&c));
    if (px < box.x || px > box.z || py < box.y || py > box.w) continue;
    int in = -1;
    if (c.feps >= 0.0f) {
      const float fy = (float)(py - box.y);
      in = cell_screen(c, (float)(px - box.x), fmaf(c.fm[1], fy, c.fm[2]), fmaf(c.fm[4], fy, c.fm[5]),
                       fmaf(c.fm[7], fy, c.fm[8]));
    }
    if (in < 0) in = cell_inside(c, x, y) ? 1 : 0;
    if (in == 1) { cell_map(c, x, y, mx, my); break; }
  }
  return make_float2(mx, my);
}

// Four taps of one output pixel, all inside the frame, regrouped per channel.  Each source row
// contributes 6 bytes (two BGR pixels) at an arbitrary byte phase: they are read as two or three
// aligned 32-bit words and brought to phase 0 with a funnel shift, then the bytes are regrouped as
// (p00, p01, p10, p11) per channel.  off0 / off1: byte offsets of the left tap in the two rows.
__device__ __forceinline__ void gather_taps(const uint8_t* __restrict__ frame, unsigned off0, unsigned off1,
                                            uint32_t& qb, uint32_t& qg, uint32_t& qr) {
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(frame) + off0;
  const uintptr_t a1 = reinterpret_cast<uintptr_t>(frame) + off1;
  const unsigned f0 = (unsigned)(a0 & 3u), f1 = (unsigned)(a1 & 3u);
  const uint32_t* w0 = reinterpret_cast<const uint32_t*>(a0 & ~(uintptr_t)3);
  const uint32_t* w1 = reinterpret_cast<const uint32_t*>(a1 & ~(uintptr_t)3);
  const uint32_t t0 = __ldg(w0), t1 = __ldg(w0 + 1), t2 = (f0 == 3u) ? __ldg(w0 + 2) : 0u;
  const uint32_t b0 = __ldg(w1), b1 = __ldg(w1 + 1), b2 = (f1 == 3u) ? __ldg(w1 + 2) : 0u;
  // s0 = bytes 0..3 (B0 G0 R0 B1), s1 = bytes 4..7 (G1 R1 . .) of each row
  const uint32_t ts0 = __funnelshift_r(t0, t1, f0 * 8u), ts1 = __funnelshift_r(t1, t2, f0 * 8u);
  const uint32_t bs0 = __funnelshift_r(b0, b1, f1 * 8u), bs1 = __funnelshift_r(b1, b2, f1 * 8u);
  qb = __byte_perm(ts0, bs0, 0x7430);                                                      // B00 B01 B10 B11
  const uint32_t tg = __byte_perm(ts0, ts1, 0x5241), bg = __byte_perm(bs0, bs1, 0x5241);   // G0 G1 R0 R1
  qg = __byte_perm(tg, bg, 0x5410);
  qr = __byte_perm(tg, bg, 0x7632);
}

// cv2.remap's fixed-point bilinear blend with two 2-way 16x8-bit dot products per channel
// against the weight pairs (w00, w01), (w10, w11), w_rc = wy_r * wx_c <= 1024: (sum + 512) >> 10.
// Byte offsets inside one frame fit 32 bits (8K BGR = 99.5 MB).  Returns the pixel as a BGRx word.
__device__ __forceinline__ uint32_t blend_interior(const uint8_t* __restrict__ frame, int pitch, int ix, int iy,
                                                   int ax, int ay) {
  uint32_t qb, qg, qr;
  const unsigned off0 = (unsigned)(iy * pitch + ix * 3);
  gather_taps(frame, off0, off0 + (unsigned)pitch, qb, qg, qr);
  const uint32_t wxp = (uint32_t)(32 - ax) | ((uint32_t)ax << 16);
  const uint32_t wa = wxp * (uint32_t)(32 - ay), wb = wxp * (uint32_t)ay;
  const uint32_t vb = __dp2a_hi(wb, qb, __dp2a_lo(wa, qb, 512u)) >> 10;
  const uint32_t vg = __dp2a_hi(wb, qg, __dp2a_lo(wa, qg, 512u)) >> 10;
  const uint32_t vr = __dp2a_hi(wb, qr, __dp2a_lo(wa, qr, 512u)) >> 10;
  return vb | (vg << 8) | (vr << 16);
}

__device__ __forceinline__ void store_bgr4(uint8_t* dst, const uint32_t (&o)[kPix], int npx, bool aligned) {
  if (npx == kPix && aligned) {     // 4 BGRx words -> 12 contiguous bytes
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
    __stcs(d32 + 0, (o[0] & 0x00ffffffu) | (o[1] << 24));
    __stcs(d32 + 1, ((o[1] >> 8) & 0x0000ffffu) | (o[2] << 16));
    __stcs(d32 + 2, ((o[2] >> 16) & 0x000000ffu) | (o[3] << 8));
  } else {
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      if (j < npx) {
        dst[3 * j + 0] = (uint8_t)(o[j] & 0xffu);
        dst[3 * j + 1] = (uint8_t)((o[j] >> 8) & 0xffu);
        dst[3 * j + 2] = (uint8_t)((o[j] >> 16) & 0xffu);
      }
    }
  }
}

// kFull: the frame is a whole number of 128 x 8 tiles (true at every standard resolution), so no
// thread needs an edge predicate.
#ifndef MF_WARP_MINBLOCKS
#define MF_WARP_MINBLOCKS 8
#endif
// No shared memory, no barriers: the candidates' parameters are read through L1 with warp-uniform
// (broadcast) 128-bit loads, so the only thing a warp ever waits for is its own data.
// kBoundsOnly: evaluate the maps of the tiles that hold a border cell and fold them into the crop-edge
// searches, touching no pixel (pass A of the streamed schedule: the crop rectangle of the whole video
// is known before the first frame has been uploaded).
template <bool kWriteMaps, bool kFull, bool kBoundsOnly>
__global__ void __launch_bounds__(kWarpThreads, MF_WARP_MINBLOCKS) warp_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out,
    const Cell* __restrict__ cells, const int* __restrict__ tile_count,
    const uint16_t* __restrict__ tile_list, int32_t* __restrict__ crop_out, float* __restrict__ map_out,
    int W, int H, int ncell, int tiles_x, int tiles_y, int bb, int bg, int br) {
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const size_t tile = (size_t)f * tiles_x * tiles_y + (size_t)blockIdx.y * tiles_x + blockIdx.x;
  const int tid = threadIdx.x;
  const Cell* fcells = cells + (size_t)f * ncell;
  const int craw = __ldg(tile_count + tile);
  if (kBoundsOnly && !(craw & kEdgeFlag)) return;
  const int nraw = craw & kCountMask;
  const bool overflow = nraw > kTileCap;
  const uint16_t* list = overflow ? nullptr : tile_list + tile * kTileCap;   // sorted by descending id
  const int ncand = overflow ? ncell : nraw;

  // thread -> pixels: warp (wx, wy) covers 32 x 4 pixels, lane = 8 columns-of-4 x 4 rows
  const int warp = tid >> 5, lane = tid & 31;
  const int py = y0 + (warp >> 2) * 4 + (lane >> 3);
  const int px0 = x0 + (warp & 3) * 32 + (lane & 7) * kPix;
  const bool active = kFull || (py < H && px0 < W);
  const int npx = kFull ? kPix : (active ? min(kPix, W - px0) : 0);
  const uint8_t* src = frames_in + (size_t)f * H * W * 3;

  float mx[kPix], my[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) { mx[j] = (float)(W + 1); my[j] = (float)(H + 1); }   // mfs.py:983-984

#if defined(MF_EXPERIMENT_NO_MAP)      // timing experiment only: cost of everything except cell search + map
  if (active) {
#pragma unroll
    for (int j = 0; j < kPix; ++j) { mx[j] = (float)(px0 + j) + 3.3f; my[j] = (float)py + 2.2f; }
  }
  if (false) {
#else
  if (active) {
#endif
    // group decision: the first candidate (descending id) that certainly contains both end pixels
    // contains the whole group; candidates both end pixels are certainly beyond on one side are skipped
    const Cell* hit = nullptr;
    bool general = overflow || npx < kPix;
    for (int k = 0; k < ncand && hit == nullptr && !general; ++k) {
      const Cell* c = fcells + __ldg(list + k);
      const int4 box = __ldg(reinterpret_cast<const int4*>(c));
      if (py < box.y || py > box.w || px0 > box.z || px0 + kPix - 1 < box.x) continue;
      const float4* cf = reinterpret_cast<const float4*>(c) + 1;
      const float4 f3 = __ldg(cf + 3);                 // fhi_y, feps
      if (f3.y < 0.0f || px0 < box.x || px0 + kPix - 1 > box.z) { general = true; break; }
      const float4 f0 = __ldg(cf), f1 = __ldg(cf + 1), f2 = __ldg(cf + 2);   // fm0..3 | fm4..7 | fm8 flo_x fhi_x flo_y
      const float fy = (float)(py - box.y);
      const float rbx = fmaf(f0.y, fy, f0.z), rby = fmaf(f1.x, fy, f1.y), rbw = fmaf(f1.w, fy, f2.x);
      const unsigned a = screen_sides(f0.x, f0.w, f1.z, rbx, rby, rbw, (float)(px0 - box.x), f2.y, f2.z, f2.w, f3.x, f3.y);
      const unsigned b = screen_sides(f0.x, f0.w, f1.z, rbx, rby, rbw, (float)(px0 + kPix - 1 - box.x), f2.y, f2.z,
                                      f2.w, f3.x, f3.y);
      if (a & b & 1u) hit = c;
      else if (!(a & b & 30u)) general = true;
    }
    if (general) {
#pragma unroll
      for (int j = 0; j < kPix; ++j)
        if (j < npx) {
          const float2 m = resolve_pixel(fcells, list, ncand, px0 + j, py, mx[j], my[j]);
          mx[j] = m.x; my[j] = m.y;
        }
    } else if (hit != nullptr) {
      const double2* hs = reinterpret_cast<const double2*>(hit->Hsu);
      const double2 h01 = __ldg(hs), h23 = __ldg(hs + 1), h45 = __ldg(hs + 2), h67 = __ldg(hs + 3);
      const double y = (double)py;
      const double yh1 = MF_MUL(y, h01.y), yh4 = MF_MUL(y, h45.x), yh7 = MF_MUL(y, h67.y);
#pragma unroll
      for (int j = 0; j < kPix; ++j)
        map_row(h01.x, h23.x, h23.y, h45.y, h67.x, (double)(px0 + j), yh1, yh4, yh7, mx[j], my[j]);
    }
  }

  int e_left = -1, e_top = -1, e_right = MF_INT_MAX, e_bottom = MF_INT_MAX;
  if (active) {
    if (kWriteMaps) {
#pragma unroll
      for (int j = 0; j < kPix; ++j)
        if (j < npx) reinterpret_cast<float2*>(map_out)[((size_t)f * H + py) * W + px0 + j] = make_float2(mx[j], my[j]);
    }
    // crop edges (mfs.py:1075-1098): |m - e| < 1 on the float32-valued map; cheap group pre-test first
    const float lo_x = fminf(fminf(mx[0], mx[1]), fminf(mx[2], mx[3])), hi_x = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    const float lo_y = fminf(fminf(my[0], my[1]), fminf(my[2], my[3])), hi_y = fmaxf(fmaxf(my[0], my[1]), fmaxf(my[2], my[3]));
    if (lo_x < 1.0f || hi_x > (float)(W - 2) || lo_y < 1.0f || hi_y > (float)(H - 2)) {
#pragma unroll
      for (int j = 0; j < kPix; ++j) {
        if (j < npx) {
          if (mx[j] > -1.0f && mx[j] < 1.0f) e_left = max(e_left, px0 + j);
          if (mx[j] > (float)(W - 2) && mx[j] < (float)W) e_right = min(e_right, px0 + j);
          if (my[j] > -1.0f && my[j] < 1.0f) e_top = max(e_top, py);
          if (my[j] > (float)(H - 2) && my[j] < (float)H) e_bottom = min(e_bottom, py);
        }
      }
    }
  }
  if (active && !kBoundsOnly) {
    // 1/32-px source coordinates and the fixed-point blend (mfs.py:1063-1069)
    const uint32_t border = (uint32_t)bb | ((uint32_t)bg << 8) | ((uint32_t)br << 16);
    uint32_t o[kPix];
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      int ix, iy, ax, ay;
      remap_coords_finite(mx[j], my[j], ix, iy, ax, ay);
#if defined(MF_EXPERIMENT_NO_GATHER)   // timing experiment only: cost of everything except the tap gathers
      if (true) { o[j] = (uint32_t)(ix * 7 + iy * 3 + ax + ay) & 0x00ffffffu; } else
#endif
      if ((unsigned)ix < (unsigned)(W - 3) && (unsigned)iy < (unsigned)(H - 1)) {
        o[j] = blend_interior(src, W * 3, ix, iy, ax, ay);
      } else if (ix < -1 || ix >= W || iy < -1 || iy >= H) {
        o[j] = border;                       // no tap inside the frame: exactly the border colour
      } else {
        uint8_t t3[3];
        remap_pixel(src, W, H, ix, iy, ax, ay, bb, bg, br, t3);
        o[j] = (uint32_t)t3[0] | ((uint32_t)t3[1] << 8) | ((uint32_t)t3[2] << 16);
      }
    }
    uint8_t* dst = frames_out + ((size_t)f * H * W + (size_t)py * W + px0) * 3;
    store_bgr4(dst, o, npx, (W & 3) == 0);
  }

  // crop edges: warp reduction, then at most four atomics per warp
  const unsigned full = 0xffffffffu;
  const bool any_edge = e_left >= 0 || e_top >= 0 || e_right != MF_INT_MAX || e_bottom != MF_INT_MAX;
  if (__any_sync(full, any_edge)) {
    e_left = __reduce_max_sync(full, e_left);
    e_top = __reduce_max_sync(full, e_top);
    e_right = __reduce_min_sync(full, e_right);
    e_bottom = __reduce_min_sync(full, e_bottom);
    if (lane == 0) {
      int32_t* cr = crop_out + 4 * f;
      if (e_left >= 0) atomicMax(cr + 0, e_left);
      if (e_top >= 0) atomicMax(cr + 1, e_top);
      if (e_right != MF_INT_MAX) atomicMin(cr + 2, e_right);
      if (e_bottom != MF_INT_MAX) atomicMin(cr + 3, e_bottom);
    }
  }
}

// ---- crop + resize -----------------------------------------------------------------------------
// crop rectangle as the device sees it: [left, top, -right, -bottom] so that ONE max-reduction
// (warp shuffle here, ncclMax across GPUs) combines per-frame / per-GPU results (mfs.py:1103-1106)
__global__ void __launch_bounds__(256) crop_combine_kernel(const int32_t* __restrict__ per_frame, int nf,
                                                           int32_t* __restrict__ enc4) {
  int v[4] = {MF_INT_MIN, MF_INT_MIN, MF_INT_MIN, MF_INT_MIN};
  for (int f = threadIdx.x; f < nf; f += blockDim.x) {
    v[0] = max(v[0], per_frame[4 * f + 0]);
    v[1] = max(v[1], per_frame[4 * f + 1]);
    v[2] = max(v[2], -per_frame[4 * f + 2]);
    v[3] = max(v[3], -per_frame[4 * f + 3]);
  }
  __shared__ int red[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int m = __reduce_max_sync(0xffffffffu, v[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = m;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int m = red[threadIdx.x][0];
    for (int w = 1; w < 8; ++w) m = max(m, red[threadIdx.x][w]);
    enc4[threadIdx.x] = m;
  }
}

__device__ __forceinline__ bool decode_crop(const int32_t* enc4, int W, int H, int& l, int& t, int& r, int& b) {
  l = enc4[0]; t = enc4[1]; r = -enc4[2]; b = -enc4[3];
  return 0 <= l && l <= r && r < W && 0 <= t && t <= b && b < H;
}

__global__ void __launch_bounds__(128) resize_table_kernel(int W, int H, const int32_t* __restrict__ enc4,
                                                           int4* __restrict__ xtab, int4* __restrict__ ytab) {
  int l, t, r, b;
  if (!decode_crop(enc4, W, H, l, t, r, b)) return;   // host raises once it reads the rectangle back
  const int sw = r - l + 1, sh = b - t + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) {
    int4 q;
    resize_coef(i, sw, W, true, q.x, q.y, q.z, q.w);
    xtab[i] = q;
  } else if (i < W + H) {
    int4 q;
    resize_coef(i - W, sh, H, false, q.x, q.y, q.z, q.w);
    ytab[i - W] = q;
  }
}

// cv2.resize of the crop window back to W x H (mfs.py:1150-1155).  Same thread layout
// as the warp kernel (4 adjacent output pixels per thread); taps come from gather_taps when the two
// columns are adjacent and away from the frame's last columns, else byte by byte.
__global__ void __launch_bounds__(128) crop_resize_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out, int W, int H,
    const int32_t* __restrict__ enc4, const int4* __restrict__ xtab, const int4* __restrict__ ytab) {
  int left, top, right_unused, bottom_unused;
  if (!decode_crop(enc4, W, H, left, top, right_unused, bottom_unused)) return;
  const int f = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int px0 = blockIdx.x * kTileW + warp * 32 + (lane & 7) * kPix;
  if (px0 >= W) return;
  const int npx = min(kPix, W - px0);
  const uint8_t* src = frames_in + (size_t)f * H * W * 3;
  const int pitch = W * 3;
  // column taps of this thread's four pixels, shared by both of its rows
  int c0[kPix], c1[kPix];
  uint32_t wx[kPix];
  int a0[kPix], a1[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    const int4 xt = __ldg(xtab + min(px0 + j, W - 1));
    c0[j] = left + xt.x; c1[j] = left + xt.y; a0[j] = xt.z; a1[j] = xt.w;
    wx[j] = (uint32_t)xt.z | ((uint32_t)xt.w << 16);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int py = blockIdx.y * kTileH + half * 4 + (lane >> 3);
    if (py >= H) continue;
    const int4 yt = __ldg(ytab + py);
    const unsigned row0 = (unsigned)((top + yt.x) * pitch), row1 = (unsigned)((top + yt.y) * pitch);
    uint32_t o[kPix];
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      o[j] = 0u;
      if (j >= npx) continue;
      uint32_t acc = 0u;
      if (c1[j] == c0[j] + 1 && c0[j] + 3 < W) {
        uint32_t q[3];
        gather_taps(src, row0 + (unsigned)(c0[j] * 3), row1 + (unsigned)(c0[j] * 3), q[0], q[1], q[2]);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const int s0 = (int)__dp2a_lo(wx[j], q[ch], 0u), s1 = (int)__dp2a_hi(wx[j], q[ch], 0u);
          // a0 + a1 <= 2049 and b0 + b1 <= 2049 bound the result by 255: no clamp needed
          const int v = (((yt.z * (s0 >> 4)) >> 16) + ((yt.w * (s1 >> 4)) >> 16) + 2) >> 2;
          acc |= (uint32_t)v << (8 * ch);
        }
      } else {
        const uint8_t* r0 = src + row0;
        const uint8_t* r1 = src + row1;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          acc |= (uint32_t)resize_blend(__ldg(r0 + c0[j] * 3 + ch), __ldg(r0 + c1[j] * 3 + ch),
                                        __ldg(r1 + c0[j] * 3 + ch), __ldg(r1 + c1[j] * 3 + ch), a0[j], a1[j],
                                        yt.z, yt.w) << (8 * ch);
      }
      o[j] = acc;
    }
    uint8_t* dst = frames_out + ((size_t)f * H * W + (size_t)py * W + px0) * 3;
    store_bgr4(dst, o, npx, (W & 3) == 0);
  }
}

// cv2.resize with the horizontal pass shared between output rows.  The crop is stretched (scale <= 1),
// so consecutive output rows read the same source row most of the time: a thread owns 4 adjacent output
// columns over kResizeRows rows and keeps the horizontal sums (a0*p0 + a1*p1) >> 4 of the two source
// rows it last used -- 12 values each -- recomputing a row only when the row index changes (a
// warp-uniform decision).  Column taps, byte phases and weights are fixed per thread, outside the row
// loop.  Vertical pass: ((b0*s0) >> 16) + ((b1*s1) >> 16) as two high multiplies by b << 16.
// Threads whose columns are clamped at the crop border (c1 != c0 + 1), partial groups and unaligned
// frames take the general per-pixel form.
#ifndef MF_RESIZE_ROWS
#define MF_RESIZE_ROWS 16
#endif
static constexpr int kResizeRows = MF_RESIZE_ROWS;

struct RowSums { uint32_t v[kPix][3]; };

__device__ __forceinline__ void resize_hsum(const uint8_t* __restrict__ row, const unsigned (&woff)[kPix],
                                            const unsigned (&shift)[kPix], const uint32_t (&wx)[kPix], RowSums& out) {
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(row + woff[j]);
    const uint32_t t0 = __ldg(w), t1 = __ldg(w + 1), t2 = shift[j] == 24u ? __ldg(w + 2) : 0u;
    const uint32_t u0 = __funnelshift_r(t0, t1, shift[j]), u1 = __funnelshift_r(t1, t2, shift[j]);   // B0 G0 R0 B1 | G1 R1 . .
    const uint32_t bg = __byte_perm(u0, u1, 0x4130), rr = __byte_perm(u0, u1, 0x0052);
    out.v[j][0] = __dp2a_lo(wx[j], bg, 0u) >> 4;
    out.v[j][1] = __dp2a_hi(wx[j], bg, 0u) >> 4;
    out.v[j][2] = __dp2a_lo(wx[j], rr, 0u) >> 4;
  }
}

#ifndef MF_RESIZE_MINBLOCKS
#define MF_RESIZE_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kWarpThreads, MF_RESIZE_MINBLOCKS) crop_resize_rows_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out, int W, int H,
    const int32_t* __restrict__ enc4, const int4* __restrict__ xtab, const int4* __restrict__ ytab) {
  int left, top, right_unused, bottom_unused;
  if (!decode_crop(enc4, W, H, left, top, right_unused, bottom_unused)) return;
  const int f = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int px0 = blockIdx.x * kTileW + lane * kPix;
  const int y_first = (blockIdx.y * (kWarpThreads / 32) + warp) * kResizeRows;
  if (px0 >= W || y_first >= H) return;
  const int npx = min(kPix, W - px0);
  const int y_end = min(H, y_first + kResizeRows);
  const uint8_t* src = frames_in + (size_t)f * H * W * 3;
  uint8_t* drow = frames_out + ((size_t)f * H * W + (size_t)y_first * W + px0) * 3;
  const unsigned pitch = (unsigned)W * 3u;
  int c0[kPix], c1[kPix], a0[kPix], a1[kPix];
  uint32_t wx[kPix];
  unsigned woff[kPix], shift[kPix];
  bool regular = npx == kPix && (pitch & 3u) == 0u && (reinterpret_cast<uintptr_t>(src) & 3u) == 0u &&
                 (reinterpret_cast<uintptr_t>(drow) & 3u) == 0u;
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    const int4 xt = __ldg(xtab + min(px0 + j, W - 1));
    c0[j] = left + xt.x; c1[j] = left + xt.y; a0[j] = xt.z; a1[j] = xt.w;
    wx[j] = (uint32_t)xt.z | ((uint32_t)xt.w << 16);
    const unsigned b = (unsigned)c0[j] * 3u;
    woff[j] = b & ~3u; shift[j] = (b & 3u) * 8u;
    regular = regular && c1[j] == c0[j] + 1 && c0[j] + 4 <= W - 1;     // 12-byte reads stay inside the row
  }
  if (regular) {
    // Two row buffers whose roles (top / bottom) swap instead of being copied when the window moves down by
    // one source row: `swapped` says which is which, and the step is instantiated once per role assignment.
    RowSums P, Q;
    int rp = -1, rq = -1;
    bool swapped = false;                                    // false: P is the top row, Q the bottom row
    auto vertical = [&](const RowSums& T, const RowSums& M, const int4& yt) {
      const uint32_t b0s = (uint32_t)yt.z << 16, b1s = (uint32_t)yt.w << 16;
      uint32_t v[kPix][3];
#pragma unroll
      for (int j = 0; j < kPix; ++j)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)   // a0 + a1 <= 2049 and b0 + b1 <= 2049 bound the result by 255: no clamp needed
          v[j][ch] = (__umulhi(b0s, T.v[j][ch]) + __umulhi(b1s, M.v[j][ch]) + 2u) >> 2;
      uint32_t* d32 = reinterpret_cast<uint32_t*>(drow);
      __stcs(d32 + 0, __byte_perm(__byte_perm(v[0][0], v[0][1], 0x0040), __byte_perm(v[0][2], v[1][0], 0x0040), 0x5410));
      __stcs(d32 + 1, __byte_perm(__byte_perm(v[1][1], v[1][2], 0x0040), __byte_perm(v[2][0], v[2][1], 0x0040), 0x5410));
      __stcs(d32 + 2, __byte_perm(__byte_perm(v[2][2], v[3][0], 0x0040), __byte_perm(v[3][1], v[3][2], 0x0040), 0x5410));
    };
    // T holds source row rt in the top role, M holds rm in the bottom role; returns true when the roles swapped
    auto step = [&](RowSums& T, RowSums& M, int& rt, int& rm, int r0, int r1, const int4& yt) -> bool {
      if (r0 != rt && r0 == rm && r1 != rm) {                // the usual move: old bottom becomes top, one new row
        resize_hsum(src + (size_t)r1 * pitch, woff, shift, wx, T);
        rt = r1;
        vertical(M, T, yt);
        return true;
      }
      if (r0 != rt) {
        if (r0 == rm) T = M; else resize_hsum(src + (size_t)r0 * pitch, woff, shift, wx, T);
        rt = r0;
      }
      if (r1 != rm) {
        if (r1 == rt) M = T; else resize_hsum(src + (size_t)r1 * pitch, woff, shift, wx, M);
        rm = r1;
      }
      vertical(T, M, yt);
      return false;
    };
    for (int py = y_first; py < y_end; ++py, drow += pitch) {
      const int4 yt = __ldg(ytab + py);
      const int r0 = top + yt.x, r1 = top + yt.y;            // warp-uniform
      const bool flip = swapped ? step(Q, P, rq, rp, r0, r1, yt) : step(P, Q, rp, rq, r0, r1, yt);
      swapped = swapped != flip;
    }
    return;
  }
  for (int py = y_first; py < y_end; ++py, drow += pitch) {   // general form, pixel by pixel
    const int4 yt = __ldg(ytab + py);
    const uint8_t* r0 = src + (size_t)(top + yt.x) * pitch;
    const uint8_t* r1 = src + (size_t)(top + yt.y) * pitch;
    for (int j = 0; j < npx; ++j)
      for (int ch = 0; ch < 3; ++ch)
        drow[3 * j + ch] = (uint8_t)resize_blend(__ldg(r0 + c0[j] * 3 + ch), __ldg(r0 + c1[j] * 3 + ch),
                                                 __ldg(r1 + c0[j] * 3 + ch), __ldg(r1 + c1[j] * 3 + ch), a0[j], a1[j],
                                                 yt.z, yt.w);
  }
}

struct WarpWorkspace {
  Cell* cells;
  int* tile_count;
  uint16_t* tile_list;
  CellFast* fast;        // fast path (warp_fast.cuh)
  CellSpan* spans;
  uint32_t* rowseg;
  uint4* lane_owner;     // [nf][H][tiles_x] x 32 uint16: owner of each lane's group of four pixels
  uint32_t* span_tab;    // [nf][R*C][span_rows]: member interval of a cell on each row of its box
  double* homs;          // [nf][R*C][16]: the two 4-point homographies of every cell (Hus, Hsu)
  int* edge_count;       // [nf]: tiles of the frame that hold a border cell
  uint16_t* edge_tiles;  // [nf][tiles]: their indices (order of arrival)
  int span_rows;
  int segcap;
};

// Rows of the span table per cell: twice the rest height plus slack covers any sane box; taller boxes are
// resolved pixel by pixel.
static int span_rows_for(int H, int R) { const int h = (H + R - 1) / R; const int s = 2 * h + 32; return s < H ? s : H; }

// Segments per 128-px tile row: wide cells (16 x 16 mesh at >= 720p) need few; 16 covers 64 x 64 meshes at 720p.
static int seg_capacity(int W, int C) { return (W / C >= 48) ? 8 : kSegMax; }

static bool carve_warp(Carver& cv, int nf, int W, int H, int R, int C, WarpWorkspace& w) {
  const size_t tiles_x = (size_t)((W + kTileW - 1) / kTileW);
  const size_t tiles = tiles_x * ((H + kTileH - 1) / kTileH);
  w.cells = cv.take<Cell>((size_t)nf * R * C);
  w.tile_count = cv.take<int>((size_t)nf * tiles);
  w.tile_list = cv.take<uint16_t>((size_t)nf * tiles * kTileCap);
  w.fast = cv.take<CellFast>((size_t)nf * R * C);
  w.spans = cv.take<CellSpan>((size_t)nf * R * C);
  w.segcap = seg_capacity(W, C);
  w.rowseg = cv.take<uint32_t>((size_t)nf * H * tiles_x * w.segcap);
  w.lane_owner = cv.take<uint4>((size_t)nf * H * tiles_x * 4);
  w.span_rows = span_rows_for(H, R);
  w.span_tab = cv.take<uint32_t>((size_t)nf * R * C * w.span_rows);
  w.homs = cv.take<double>((size_t)nf * R * C * 16);
  w.edge_count = cv.take<int>((size_t)nf);
  w.edge_tiles = cv.take<uint16_t>((size_t)nf * tiles);
  return cv.ok();
}

}  // namespace mf

#include "warp_fast.cuh"

// The fast path needs cell ids below the two reserved segment owners and frames wide enough for its
// 5-word row reads; MF_WARP_GENERIC=1 forces the generic kernel (A/B comparisons).
static bool use_fast_path(int W, int H, int R, int C) {
  static const bool forced_generic = [] { const char* e = getenv("MF_WARP_GENERIC"); return e && e[0] == '1'; }();
  const int64_t tiles = (int64_t)((W + mf::kTileW - 1) / mf::kTileW) * ((H + mf::kTileH - 1) / mf::kTileH);
  return !forced_generic && W >= 16 && H >= 2 && W <= 32767 && H <= 32767 && (int64_t)R * C < (int64_t)mf::kSegStraddle &&
         tiles <= 65535;
}

extern "C" size_t mf_warp_workspace_bytes(int nf, int W, int H, int R, int C) {
  if (nf <= 0 || W <= 0 || H <= 0 || R <= 0 || C <= 0) return 0;
  mf::Carver cv(nullptr, 0);
  mf::WarpWorkspace w;
  mf::carve_warp(cv, nf, W, H, R, C, w);
  return mf::align_up(cv.used, 256);
}

static mf::WarpTables tables_of(const mf::WarpWorkspace& w, int R, int C, int tiles_x, int tiles_y) {
  mf::WarpTables t;
  t.cells = w.cells; t.fast = w.fast; t.tile_count = w.tile_count; t.tile_list = w.tile_list; t.rowseg = w.rowseg;
  t.lane_owner = (const uint16_t*)w.lane_owner; t.segcap = w.segcap; t.ncell = R * C; t.tiles_x = tiles_x; t.tiles_y = tiles_y;
  return t;
}

// memset + cell_setup + tile_sort (+ cell_spans + row_segments on the fast path): everything that depends on the
// vertex paths only.  On the fast path row_segments also folds the crop edges of every frame into crop_out.
static int prepare_cells(const double* u, const double* s, const float* vertex_xy, int nf, int W, int H, int R,
                         int C, int32_t* crop_out, const mf::WarpWorkspace& w, bool fast, cudaStream_t st) {
  const int tiles_x = (W + mf::kTileW - 1) / mf::kTileW, tiles_y = (H + mf::kTileH - 1) / mf::kTileH;
  cudaError_t ce = cudaMemsetAsync(w.tile_count, 0, (size_t)nf * tiles_x * tiles_y * sizeof(int), st);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(w.edge_count, 0, (size_t)nf * sizeof(int), st);
  if (ce != cudaSuccess) return mf::fail(MF_E_LAUNCH, "warp: memset: %s", cudaGetErrorString(ce));
  const int64_t ncells = (int64_t)nf * R * C;
  mf::cell_homographies_kernel<<<(unsigned)((ncells * 16 + 127) / 128), 128, 0, st>>>(u, s, vertex_xy, nf, R, C, w.homs);
  if (int e = mf::check_launch("cell_homographies")) return e;
  mf::cell_setup_kernel<<<(unsigned)((ncells + 127) / 128), 128, 0, st>>>(
      u, s, vertex_xy, w.homs, nf, W, H, R, C, tiles_x, tiles_y, w.cells, fast ? w.fast : nullptr, fast ? w.spans : nullptr,
      w.tile_count, w.tile_list, crop_out);
  if (int e = mf::check_launch("cell_setup")) return e;
  const int64_t ntiles = (int64_t)nf * tiles_x * tiles_y;
  mf::tile_sort_kernel<<<(unsigned)((ntiles + 127) / 128), 128, 0, st>>>(w.tile_count, w.tile_list, ntiles, tiles_x * tiles_y,
                                                                       w.edge_count, w.edge_tiles);
  if (int e = mf::check_launch("tile_sort")) return e;
  if (fast) {
    const int64_t nspan = ncells * mf::kSpanChunks;
    mf::cell_spans_kernel<<<(unsigned)((nspan + 127) / 128), 128, 0, st>>>(w.cells, w.spans, ncells, w.span_rows, w.span_tab);
    if (int e = mf::check_launch("cell_spans")) return e;
    const int64_t nrows = (int64_t)nf * H * tiles_x;
    if (w.segcap == 8)
      mf::row_segments_kernel<8><<<(unsigned)((nrows + 127) / 128), 128, 0, st>>>(
          w.cells, w.span_tab, w.span_rows, w.tile_count, w.tile_list, nf, W, H, R * C, tiles_x, tiles_y, w.rowseg, w.lane_owner);
    else
      mf::row_segments_kernel<mf::kSegMax><<<(unsigned)((nrows + 127) / 128), 128, 0, st>>>(
          w.cells, w.span_tab, w.span_rows, w.tile_count, w.tile_list, nf, W, H, R * C, tiles_x, tiles_y, w.rowseg, w.lane_owner);
    if (int e = mf::check_launch("row_segments")) return e;
    // one thread per (frame, listed border tile, row of the tile); a CTA strides over its frame's list
    const int all_blocks = (tiles_x * tiles_y + mf::kCropEdgeSlotsPerCta - 1) / mf::kCropEdgeSlotsPerCta;
    const int slot_blocks = all_blocks < mf::kCropEdgeSlotBlocks ? all_blocks : mf::kCropEdgeSlotBlocks;
    mf::crop_edges_kernel<<<dim3((unsigned)slot_blocks, (unsigned)nf), 128, 0, st>>>(
        w.cells, w.tile_count, w.tile_list, w.rowseg, w.edge_count, w.edge_tiles, nf, W, H, R * C, tiles_x, tiles_y, w.segcap,
        crop_out);
    return mf::check_launch("crop_edges");
  }
  return MF_OK;
}

static dim3 fast_grid(int nf, int W, int H) {
  return dim3((unsigned)((W + mf::kTileW - 1) / mf::kTileW), (unsigned)((H + mf::kFastTileH - 1) / mf::kFastTileH),
              (unsigned)nf);
}

extern "C" int mf_warp_prepare(const double* u, const double* s, const float* vertex_xy, int nf, int W,
                               int H, int R, int C, int32_t* crop_out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  MF_REQUIRE(u && s && vertex_xy && crop_out && workspace, "mf_warp_prepare: null pointer");
  MF_REQUIRE(nf > 0 && nf <= 65535 && W > 1 && H > 1 && R > 0 && C > 0 && R * C <= 65535,
             "mf_warp_prepare: bad sizes");
  mf::Carver cv(workspace, workspace_bytes);
  mf::WarpWorkspace w;
  if (!mf::carve_warp(cv, nf, W, H, R, C, w))
    return mf::fail(MF_E_WORKSPACE, "mf_warp_prepare: workspace %zu < %zu bytes", workspace_bytes, cv.used);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast = use_fast_path(W, H, R, C);
  if (int e = prepare_cells(u, s, vertex_xy, nf, W, H, R, C, crop_out, w, fast, st)) return e;
  if (fast) return MF_OK;
  // frames too small for the row-segment tables: the generic kernel evaluates the maps of the border tiles
  const int tiles_x = (W + mf::kTileW - 1) / mf::kTileW, tiles_y = (H + mf::kTileH - 1) / mf::kTileH;
  const dim3 grid((unsigned)tiles_x, (unsigned)tiles_y, (unsigned)nf);
  if ((W % mf::kTileW == 0) && (H % mf::kTileH == 0))
    mf::warp_kernel<false, true, true><<<grid, mf::kWarpThreads, 0, st>>>(
        nullptr, nullptr, w.cells, w.tile_count, w.tile_list, crop_out, nullptr, W, H, R * C, tiles_x, tiles_y, 0, 0, 0);
  else
    mf::warp_kernel<false, false, true><<<grid, mf::kWarpThreads, 0, st>>>(
        nullptr, nullptr, w.cells, w.tile_count, w.tile_list, crop_out, nullptr, W, H, R * C, tiles_x, tiles_y, 0, 0, 0);
  return mf::check_launch("warp_bounds");
}

extern "C" int mf_warp_crop_bounds(const double* u, const double* s, const float* vertex_xy, int nf, int W,
                                   int H, int R, int C, int32_t* crop_out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  return mf_warp_prepare(u, s, vertex_xy, nf, W, H, R, C, crop_out, workspace, workspace_bytes, stream);
}

extern "C" int mf_warp_frames(const uint8_t* frames_in, const double* u, const double* s,
                              const float* vertex_xy, int nf, int W, int H, int R, int C, int border_b,
                              int border_g, int border_r, uint8_t* frames_out, int32_t* crop_out,
                              float* map_out, void* workspace, size_t workspace_bytes, void* stream) {
  MF_REQUIRE(frames_in && u && s && vertex_xy && frames_out && crop_out && workspace,
             "mf_warp_frames: null pointer");
  MF_REQUIRE(nf > 0 && W > 1 && H > 1 && R > 0 && C > 0, "mf_warp_frames: bad sizes");
  MF_REQUIRE(R * C <= 65535, "mf_warp_frames: at most 65535 mesh cells");
  MF_REQUIRE(nf <= 65535, "mf_warp_frames: at most 65535 frames per call");
  MF_REQUIRE(frames_in != frames_out, "mf_warp_frames: in-place warp is not possible");
  mf::Carver cv(workspace, workspace_bytes);
  mf::WarpWorkspace w;
  if (!mf::carve_warp(cv, nf, W, H, R, C, w))
    return mf::fail(MF_E_WORKSPACE, "mf_warp_frames: workspace %zu < %zu bytes", workspace_bytes, cv.used);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = (W + mf::kTileW - 1) / mf::kTileW, tiles_y = (H + mf::kTileH - 1) / mf::kTileH;
  // the float32 maps (map_out, parity checks only) exist in the generic kernel alone
  const bool fast = map_out == nullptr && use_fast_path(W, H, R, C);
  if (int e = prepare_cells(u, s, vertex_xy, nf, W, H, R, C, crop_out, w, fast, st)) return e;
  if (fast) {
    const uint32_t border = (uint32_t)(border_b & 255) | ((uint32_t)(border_g & 255) << 8) | ((uint32_t)(border_r & 255) << 16);
    mf::warp_fast_kernel<<<fast_grid(nf, W, H), mf::kWarpThreads, 0, st>>>(frames_in, frames_out,
                                                                           tables_of(w, R, C, tiles_x, tiles_y), W, H, border);
    return mf::check_launch("warp_fast");
  }
  const dim3 grid((unsigned)tiles_x, (unsigned)tiles_y, (unsigned)nf);
  const bool full = (W % mf::kTileW == 0) && (H % mf::kTileH == 0);
#define MF_LAUNCH_WARP(MAPS, FULL)                                                                        \
  mf::warp_kernel<MAPS, FULL, false><<<grid, mf::kWarpThreads, 0, st>>>(                                 \
      frames_in, frames_out, w.cells, w.tile_count, w.tile_list, crop_out, map_out, W, H, R * C, tiles_x, \
      tiles_y, border_b, border_g, border_r)
  if (map_out) { if (full) MF_LAUNCH_WARP(true, true); else MF_LAUNCH_WARP(true, false); }
  else { if (full) MF_LAUNCH_WARP(false, true); else MF_LAUNCH_WARP(false, false); }
#undef MF_LAUNCH_WARP
  return mf::check_launch("warp");
}

extern "C" size_t mf_crop_resize_workspace_bytes(int W, int H) {
  if (W <= 0 || H <= 0) return 0;
  return 256 + mf::align_up((size_t)W * sizeof(int4), 256) + mf::align_up((size_t)H * sizeof(int4), 256);
}

static int launch_resize_tables(int W, int H, const int32_t* enc4, void* workspace, cudaStream_t st, int4*& xtab, int4*& ytab) {
  xtab = (int4*)((char*)workspace + 256);
  ytab = (int4*)((char*)workspace + 256 + mf::align_up((size_t)W * sizeof(int4), 256));
  mf::resize_table_kernel<<<(W + H + 127) / 128, 128, 0, st>>>(W, H, enc4, xtab, ytab);
  return mf::check_launch("resize_table");
}

static int launch_crop_resize(const uint8_t* frames_in, int nf, int W, int H, const int32_t* enc4,
                              uint8_t* frames_out, void* workspace, cudaStream_t st) {
  int4 *xtab, *ytab;
  if (int e = launch_resize_tables(W, H, enc4, workspace, st, xtab, ytab)) return e;
  static const bool generic = [] { const char* e = getenv("MF_RESIZE_GENERIC"); return e && e[0] == '1'; }();
  if (generic) {
    const dim3 grid((unsigned)((W + mf::kTileW - 1) / mf::kTileW), (unsigned)((H + mf::kTileH - 1) / mf::kTileH),
                    (unsigned)nf);
    mf::crop_resize_kernel<<<grid, 128, 0, st>>>(frames_in, frames_out, W, H, enc4, xtab, ytab);
    return mf::check_launch("crop_resize");
  }
  const int rows_per_cta = (mf::kWarpThreads / 32) * mf::kResizeRows;
  const dim3 grid((unsigned)((W + mf::kTileW - 1) / mf::kTileW), (unsigned)((H + rows_per_cta - 1) / rows_per_cta),
                  (unsigned)nf);
  mf::crop_resize_rows_kernel<<<grid, mf::kWarpThreads, 0, st>>>(frames_in, frames_out, W, H, enc4, xtab, ytab);
  return mf::check_launch("crop_resize_rows");
}

extern "C" int mf_crop_combine(const int32_t* per_frame_crop, int nf, int32_t* crop_enc_out, void* stream) {
  MF_REQUIRE(per_frame_crop && crop_enc_out && nf > 0, "mf_crop_combine: bad arguments");
  mf::crop_combine_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(per_frame_crop, nf, crop_enc_out);
  return mf::check_launch("crop_combine");
}

extern "C" int mf_crop_resize(const uint8_t* frames_in, int nf, int W, int H, int left, int top, int right,
                              int bottom, uint8_t* frames_out, void* workspace, size_t workspace_bytes,
                              void* stream) {
  MF_REQUIRE(frames_in && frames_out && workspace, "mf_crop_resize: null pointer");
  MF_REQUIRE(nf > 0 && W > 0 && H > 0, "mf_crop_resize: bad sizes");
  MF_REQUIRE(nf <= 65535, "mf_crop_resize: at most 65535 frames per call");
  MF_REQUIRE(0 <= left && left <= right && right < W && 0 <= top && top <= bottom && bottom < H,
             "mf_crop_resize: crop rectangle (%d,%d,%d,%d) outside the %dx%d frame or empty", left, top,
             right, bottom, W, H);
  MF_REQUIRE(frames_in != frames_out, "mf_crop_resize: in-place resize is not possible");
  if (workspace_bytes < mf_crop_resize_workspace_bytes(W, H))
    return mf::fail(MF_E_WORKSPACE, "mf_crop_resize: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t enc[4] = {left, top, -right, -bottom};
  // 16 bytes through the kernel-parameter-like path: cudaMemcpyAsync from pageable memory copies the
  // source before returning, so the stack array may die right after
  cudaError_t ce = cudaMemcpyAsync(workspace, enc, sizeof(enc), cudaMemcpyHostToDevice, st);
  if (ce != cudaSuccess) return mf::fail(MF_E_LAUNCH, "mf_crop_resize: memcpy: %s", cudaGetErrorString(ce));
  return launch_crop_resize(frames_in, nf, W, H, (const int32_t*)workspace, frames_out, workspace, st);
}

extern "C" int mf_crop_resize_device(const uint8_t* frames_in, int nf, int W, int H, const int32_t* crop_enc,
                                     uint8_t* frames_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  MF_REQUIRE(frames_in && frames_out && workspace && crop_enc, "mf_crop_resize_device: null pointer");
  MF_REQUIRE(nf > 0 && nf <= 65535 && W > 0 && H > 0, "mf_crop_resize_device: bad sizes");
  MF_REQUIRE(frames_in != frames_out, "mf_crop_resize_device: in-place resize is not possible");
  if (workspace_bytes < mf_crop_resize_workspace_bytes(W, H))
    return mf::fail(MF_E_WORKSPACE, "mf_crop_resize_device: workspace too small");
  return launch_crop_resize(frames_in, nf, W, H, crop_enc, frames_out, workspace, (cudaStream_t)stream);
}

// Fused pass B: warp + crop + resize of frames [first_frame, first_frame + nf) of the video whose tables
// mf_warp_prepare left in `workspace` (carved for table_frames frames).
extern "C" int mf_warp_resize_frames(const uint8_t* frames_in, int nf, int first_frame, int table_frames, int W, int H,
                                     int R, int C, int border_b, int border_g, int border_r, const int32_t* crop_enc,
                                     uint8_t* frames_out, void* workspace, size_t workspace_bytes,
                                     void* resize_workspace, size_t resize_workspace_bytes, void* stream) {
  MF_REQUIRE(frames_in && frames_out && crop_enc && workspace && resize_workspace, "mf_warp_resize_frames: null pointer");
  MF_REQUIRE(nf > 0 && first_frame >= 0 && table_frames > 0 && first_frame + nf <= table_frames && table_frames <= 65535,
             "mf_warp_resize_frames: frames [%d, %d) are not inside the %d prepared frames", first_frame, first_frame + nf,
             table_frames);
  MF_REQUIRE(W > 1 && H > 1 && R > 0 && C > 0 && R * C <= 65535, "mf_warp_resize_frames: bad sizes");
  MF_REQUIRE(frames_in != frames_out, "mf_warp_resize_frames: in-place operation is not possible");
  if (!use_fast_path(W, H, R, C))
    return mf::fail(MF_E_UNSUPPORTED, "mf_warp_resize_frames: no row-segment tables at %dx%d with %dx%d cells "
                                      "(use mf_warp_frames + mf_crop_resize_device)", W, H, R, C);
  if (resize_workspace_bytes < mf_crop_resize_workspace_bytes(W, H))
    return mf::fail(MF_E_WORKSPACE, "mf_warp_resize_frames: resize workspace too small");
  mf::Carver cv(workspace, workspace_bytes);
  mf::WarpWorkspace w;
  if (!mf::carve_warp(cv, table_frames, W, H, R, C, w))
    return mf::fail(MF_E_WORKSPACE, "mf_warp_resize_frames: workspace %zu < %zu bytes", workspace_bytes, cv.used);
  cudaStream_t st = (cudaStream_t)stream;
  int4 *xtab, *ytab;
  if (int e = launch_resize_tables(W, H, crop_enc, resize_workspace, st, xtab, ytab)) return e;
  const int tiles_x = (W + mf::kTileW - 1) / mf::kTileW, tiles_y = (H + mf::kTileH - 1) / mf::kTileH;
  if (sizeof(mf::FusedShared) > 48 * 1024) {                 // opt-in once per device
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !opted[dev]) {
      cudaError_t ce = cudaFuncSetAttribute(mf::warp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(mf::FusedShared));
      if (ce != cudaSuccess) return mf::fail(MF_E_LAUNCH, "mf_warp_resize_frames: shared memory opt-in: %s", cudaGetErrorString(ce));
      opted[dev] = true;
    }
  }
  const uint32_t border = (uint32_t)(border_b & 255) | ((uint32_t)(border_g & 255) << 8) | ((uint32_t)(border_r & 255) << 16);
  const dim3 grid((unsigned)((W + mf::kOutTileW - 1) / mf::kOutTileW), (unsigned)((H + mf::kOutTileH - 1) / mf::kOutTileH),
                  (unsigned)nf);
  mf::warp_fused_kernel<<<grid, mf::kWarpThreads, sizeof(mf::FusedShared), st>>>(
      frames_in, frames_out, tables_of(w, R, C, tiles_x, tiles_y), first_frame, W, H, border, crop_enc, xtab, ytab);
  return mf::check_launch("warp_fused");
}
