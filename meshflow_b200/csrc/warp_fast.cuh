// Production path of the mesh warp (included by warp.cu).
//
// Every decision that does not depend on pixel data is taken before the pixel pass, from the vertex
// paths alone (mf_warp_prepare):
//
//   cell_spans_kernel   : one thread per (frame, cell, four rows of the cell's box): exact member interval of
//                         the cell on a row (span_of_row: four half-planes + the exact test for a pixel
//                         within rounding noise of a boundary).
//   row_segments_kernel : one thread per (frame, row, 128-px tile): the candidates' intervals resolved by
//                         "the last cell written wins" into <= 16 sorted segments (first x, cell id) and the
//                         owner of each group of four pixels (what the pixel kernels read).
//   crop_edges_kernel   : the frame's crop edges (mfs.py:1075-1098) from the segments of border cells by the
//                         closed-form band search of mf_math.cuh (segment_crop_edges): the crop rectangle
//                         of a video is known before a single pixel has been read.
//
// and two pixel kernels share one row loop (warp_rows):
//
//   warp_fast_kernel    : stabilized frames to global memory (the reference's stage output,
//                         mfs.py:909-1100).  CTA = 128 x 120 output pixels, warp = 128 x 15.
//   warp_fused_kernel   : pass B of the streamed schedule.  The crop rectangle is already known, so a CTA
//                         produces the stabilized pixels of ONE 120 x 64 tile of the FINAL frame (+ the one
//                         row / column of overlap cv2.resize's taps need) into shared memory (BGRx) and resizes
//                         from there (mfs.py:1111-1157): the stabilized frame never exists in DRAM and the
//                         pixels outside the crop rectangle are never computed.
//
// warp_rows: four adjacent pixels per thread and row.  Per row a thread reads its group's owner (one
// 16-bit load, requested a row ahead), the cell's 64-byte parameter block from L1, evaluates the remap
// coordinates in float32 in box-centred form (error < eps, see mf_math.cuh) and keeps rint() only when
// the value is outside the rounding band; when the four footprints are adjacent the two source rows are
// read once as 4-5 aligned words each (the next row's new line is prefetched into L1), brought to
// phase 0 with funnel shifts, and the taps are regrouped with constant-selector PRMTs for the dp2a
// blend.  Everything else -- pixels inside the rounding band, groups that straddle a segment,
// non-adjacent footprints, taps outside the frame, irregular cells -- goes into the warp's shared-memory
// list as ONE entry per group (slot from a ballot, so a column of straddling groups costs one store per
// row, not a serial loop) and is handled by the same warp right after its rows, lanes packed:
// medium_pixel() (float32 coordinate, per-pixel tap fetch) for pixels that only failed a group
// condition, slow_pixel() (the reference's float64 sequence) for the rest.
//
// Results are identical to warp_kernel (tests/test_gpu_parity.py compares all of them with the oracle).
#pragma once

namespace mf {

// Member interval of every cell on every row of its support box.  span_tab[(f * ncell + id) * span_rows + (y - by0)] = a | b << 16 (a > b: no member pixel;
// kSpanIrregular: the row cannot be described by an interval with certainty).  Cells whose box is taller than
// span_rows are treated as irregular by row_segments_kernel.
static constexpr uint32_t kSpanIrregular = 0xffffffffu;
static constexpr uint32_t kSpanEmpty = 0x00000001u;          // a = 1, b = 0

static constexpr int kSpanRowsPerThread = 4;                 // independent rows per thread: four chains in flight

// Thread = (frame, cell, chunk of four rows).  A cell's box is rarely taller than 1.5 x the cell, far less than the
// span table's row capacity, so only kSpanChunks chunks are launched per cell and a thread strides over the box
// (round 1 launched one thread per table chunk: 40 % of the lanes had nothing to do).
static constexpr int kSpanChunks = 24;

__global__ void __launch_bounds__(128) cell_spans_kernel(const Cell* __restrict__ cells, const CellSpan* __restrict__ spans,
                                                         int64_t ncells_total, int span_rows, uint32_t* __restrict__ span_tab) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ncells_total * kSpanChunks) return;
  const int64_t cid = idx / kSpanChunks;
  const int chunk = (int)(idx - cid * kSpanChunks);
  const Cell& c = cells[cid];
  const int4 box = __ldg(reinterpret_cast<const int4*>(&c));
  if (box.x > box.z || box.y + chunk * kSpanRowsPerThread > box.w) return;
  const CellSpan& sp = spans[cid];
  const bool regular = sp.regular != 0;
  uint32_t* out = span_tab + cid * span_rows;
  for (int rl0 = chunk * kSpanRowsPerThread; rl0 < span_rows && box.y + rl0 <= box.w; rl0 += kSpanChunks * kSpanRowsPerThread) {
#pragma unroll
    for (int k = 0; k < kSpanRowsPerThread; ++k) {
      const int rl = rl0 + k, y = box.y + rl;
      if (rl >= span_rows || y > box.w) break;
      uint32_t v = kSpanIrregular;
      if (regular) {
        int a, b;
        const int st = span_of_row(c, sp, y, box.x, box.z, a, b);
        if (st == 0) v = (uint32_t)a | ((uint32_t)b << 16);
        else if (st == 1) v = kSpanEmpty;
      }
      out[rl] = v;
    }
  }
}

// Crop edges contributed by one 128-pixel tile row (only rows of tiles that hold a border cell get here).
__device__ __forceinline__ void row_crop_edges(const Cell* __restrict__ fcells, const uint16_t* __restrict__ list, int nraw,
                                            int ncell, const unsigned* seg, int ns, int x0, int x1, int y, int W, int H,
                                            int32_t* __restrict__ cr) {
  int e[4] = {cr[0], cr[1], cr[2], cr[3]};                  // stale values are fine: they only prune
  const int e0[4] = {e[0], e[1], e[2], e[3]};
  if (ns < 0) {
    // ownership of this tile row is not known per interval: every pixel the exact way
    const bool overflow = nraw > kTileCap;
    for (int px = x0; px <= x1; ++px) {
      const float2 m = resolve_pixel(fcells, overflow ? nullptr : list, overflow ? ncell : nraw, px, y, (float)(W + 1), (float)(H + 1));
      if (m.x > -1.0f && m.x < 1.0f && px > e[0]) e[0] = px;
      if (m.y > -1.0f && m.y < 1.0f && y > e[1]) e[1] = y;
      if (m.x > (float)(W - 2) && m.x < (float)W && px < e[2]) e[2] = px;
      if (m.y > (float)(H - 2) && m.y < (float)H && y < e[3]) e[3] = y;
    }
  } else {
    for (int i = 0; i < ns; ++i) {
      const unsigned id = seg[i] & 0xffffu;
      if (id == kSegNone) continue;                          // default map (W+1, H+1): never a hit
      const int xa = (int)(seg[i] >> 16), xb = i + 1 < ns ? (int)(seg[i + 1] >> 16) - 1 : x1;
      segment_crop_edges(fcells[id], xa, xb, y, W, H, e);
    }
  }
  if (e[0] > e0[0]) atomicMax(cr + 0, e[0]);
  if (e[1] > e0[1]) atomicMax(cr + 1, e[1]);
  if (e[2] < e0[2]) atomicMin(cr + 2, e[2]);
  if (e[3] < e0[3]) atomicMin(cr + 3, e[3]);
}

// packed[w] holds the 16-bit entries of lanes 2w (low half) and 2w + 1: entries of lanes >= first_lane become `pair`'s.
__device__ __forceinline__ void fill_lanes_from(uint32_t (&packed)[16], int first_lane, uint32_t pair) {
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    uint32_t m;
    asm("shl.b32 %0, %1, %2;" : "=r"(m) : "r"(0xffffffffu), "r"((unsigned)max(16 * first_lane - 32 * w, 0)));   // shl clamps: >= 32 gives 0
    packed[w] = (packed[w] & ~m) | (pair & m);
  }
}

template <int CAP>
__global__ void __launch_bounds__(128) row_segments_kernel(
    const Cell* __restrict__ cells, const uint32_t* __restrict__ span_tab, int span_rows, const int* __restrict__ tile_count,
    const uint16_t* __restrict__ tile_list, int nf, int W, int H, int ncell, int tiles_x, int tiles_y,
    uint32_t* __restrict__ rowseg, uint4* __restrict__ lane_owner) {
  static_assert(CAP == 8 || CAP == 16, "rowseg entries per tile row (seg_capacity)");
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)nf * H * tiles_x;
  if (idx >= total) return;
  const int tx = (int)(idx % tiles_x);
  const int64_t fy = idx / tiles_x;
  const int y = (int)(fy % H), f = (int)(fy / H);
  const int x0 = tx * kTileW, x1 = min(W - 1, x0 + kTileW - 1);
  const size_t tile = (size_t)f * tiles_x * tiles_y + (size_t)(y / kTileH) * tiles_x + tx;
  const int craw = __ldg(tile_count + tile);
  const int nraw = craw & kCountMask;
  SegBuilder<CAP> sb;                                        // all of it in registers
  sb.begin(x0, x1);
  bool irregular = nraw > kTileCap;
  const Cell* fcells = cells + (size_t)f * ncell;
  const uint32_t* fspan = span_tab + (size_t)f * ncell * span_rows;
  const uint16_t* list = tile_list + tile * kTileCap;       // sorted by descending id
  // candidates four at a time: ids, boxes and spans of a group are independent loads
  bool done = false;
  for (int k0 = 0; k0 < nraw && !irregular && !done; k0 += 4) {
    int id[4];
    int4 box[4];
    uint32_t spn[4];
    bool use[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) id[j] = k0 + j < nraw ? (int)__ldg(list + k0 + j) : -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) box[j] = id[j] >= 0 ? __ldg(reinterpret_cast<const int4*>(fcells + id[j])) : make_int4(1, 1, 0, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      use[j] = !(y < box[j].y || y > box[j].w || x1 < box[j].x || x0 > box[j].z);
      const bool tall = use[j] && y - box[j].y >= span_rows;
      spn[j] = tall ? kSpanIrregular : (use[j] ? __ldg(fspan + (size_t)id[j] * span_rows + (y - box[j].y)) : kSpanEmpty);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!use[j] || irregular || done) continue;
      if (spn[j] == kSpanIrregular) { irregular = true; continue; }
      const int a = max((int)(spn[j] & 0xffffu), x0), b = min((int)(spn[j] >> 16), x1);
      if (a <= b) sb.cover(a, b, (unsigned)id[j]);
      if (sb.overflow) { irregular = true; continue; }
      if (sb.done()) done = true;
    }
  }
  const int ns = irregular ? -1 : sb.finish();
  if (ns < 0) {
    sb.seg[0] = ((unsigned)x0 << 16) | kSegIrregular;
#pragma unroll
    for (int i = 1; i < CAP; ++i) sb.seg[i] = kSegSentinel;
  }
  uint4* out = reinterpret_cast<uint4*>(rowseg + (size_t)idx * CAP);
#pragma unroll
  for (int i = 0; i < CAP / 4; ++i) out[i] = make_uint4(sb.seg[4 * i], sb.seg[4 * i + 1], sb.seg[4 * i + 2], sb.seg[4 * i + 3]);
  // owner of each lane's group of four pixels (what the pixel kernel reads: one 16-bit load per lane and row).
  // Segment by segment: a segment that starts r pixels into the tile marks the group that holds its first pixel as
  // straddling (unless it starts on a group boundary) and owns every group from the next one on -- two masked fills
  // of the 16 packed words per segment start (a row-tile has 2.3 segments on average).
  uint32_t packed[16];
  if (ns < 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) packed[i] = kSegIrregular | (kSegIrregular << 16);
  } else {
    const uint32_t first = (sb.seg[0] & 0xffffu) * 0x10001u;
#pragma unroll
    for (int i = 0; i < 16; ++i) packed[i] = first;
#pragma unroll
    for (int i = 1; i < CAP; ++i) {
      if (i < ns) {
        const unsigned sg = sb.seg[i];
        const int r = (int)(sg >> 16) - x0;                  // 1 .. 127
        fill_lanes_from(packed, r >> 2, kSegStraddle * 0x10001u);
        fill_lanes_from(packed, (r + 3) >> 2, (sg & 0xffffu) * 0x10001u);
      }
    }
  }
  uint4* lo = lane_owner + (size_t)idx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) lo[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
}

// Crop edges of every frame (mfs.py:1075-1098) from the row segments of the tiles that hold a border cell (listed
// per frame by tile_sort_kernel): one thread per (frame, listed tile, row of the tile).  A kernel of its own so
// that the float64 registers of the band search do not burden the segment builder.  Grid = (kCropEdgeSlotBlocks,
// frames): a CTA takes 128 / kTileH listed tiles (all their rows) per pass and strides over the frame's list, so that
// CTAs are only launched for tiles that can be listed (a thread per row of every tile of the frame meant 75 % CTAs
// that left at once).
static constexpr int kCropEdgeSlotBlocks = 32;               // 32 x 16 = 512 listed tiles per pass; frames with more loop
static constexpr int kCropEdgeSlotsPerCta = 128 / kTileH;
static_assert(128 % kTileH == 0, "crop_edges_kernel: whole tiles per CTA");

__global__ void __launch_bounds__(128) crop_edges_kernel(
    const Cell* __restrict__ cells, const int* __restrict__ tile_count, const uint16_t* __restrict__ tile_list,
    const uint32_t* __restrict__ rowseg, const int* __restrict__ edge_count, const uint16_t* __restrict__ edge_tiles, int nf,
    int W, int H, int ncell, int tiles_x, int tiles_y, int segcap, int32_t* __restrict__ crop_out) {
  const int ntiles = tiles_x * tiles_y;
  const int f = blockIdx.y;
  const int ry = threadIdx.x % kTileH;
  const int listed = __ldg(edge_count + f);
  for (int slot = blockIdx.x * kCropEdgeSlotsPerCta + threadIdx.x / kTileH; slot < listed; slot += gridDim.x * kCropEdgeSlotsPerCta) {
    const int t = (int)__ldg(edge_tiles + (size_t)f * ntiles + slot);
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int y = ty * kTileH + ry;
    if (y >= H) continue;
    const size_t tile = (size_t)f * ntiles + t;
    const int craw = __ldg(tile_count + tile);
    const int x0 = tx * kTileW, x1 = min(W - 1, x0 + kTileW - 1);
    const uint32_t* rs = rowseg + (((size_t)f * H + y) * tiles_x + tx) * segcap;
    unsigned seg[kSegMax];                                   // the whole list in independent 16-byte loads (segcap is 8 or 16)
#pragma unroll
    for (int i = 0; i < kSegMax / 4; ++i) {
      const uint4 q = 4 * i < segcap ? __ldg(reinterpret_cast<const uint4*>(rs) + i) : make_uint4(kSegSentinel, kSegSentinel, kSegSentinel, kSegSentinel);
      seg[4 * i] = q.x; seg[4 * i + 1] = q.y; seg[4 * i + 2] = q.z; seg[4 * i + 3] = q.w;
    }
    int ns = 0;
#pragma unroll
    for (int i = 0; i < kSegMax; ++i) ns += seg[i] != kSegSentinel ? 1 : 0;    // entries are contiguous from the front
    if ((seg[0] & 0xffffu) == kSegIrregular) ns = -1;
    row_crop_edges(cells + (size_t)f * ncell, tile_list + tile * kTileCap, craw & kCountMask, ncell, seg, ns, x0, x1, y, W, H,
                   crop_out + 4 * f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pixel pass
// ---------------------------------------------------------------------------------------------------------------
struct WarpTables {
  const Cell* cells;          // [frames][ncell]
  const CellFast* fast;
  const int* tile_count;
  const uint16_t* tile_list;
  const uint32_t* rowseg;
  const uint16_t* lane_owner;
  int segcap, ncell, tiles_x, tiles_y;
};

// Where a pixel kernel puts its stabilized pixels: global memory (row pitch 3 W) or the CTA's shared-memory tile.
// `base` addresses pixel (ox, oy); both kinds are reached through generic stores on the rare paths, the hot path
// uses st.shared / st.global.cs directly.
struct PixelSink {
  uint8_t* base;
  unsigned pitch;
  int ox, oy;
  unsigned px_bytes;          // 3: packed BGR (frames in global memory); 4: BGRx (the fused kernel's shared-memory tile)
  __device__ __forceinline__ uint8_t* at(int px, int py) const {
    return base + (size_t)(py - oy) * pitch + (size_t)(px - ox) * px_bytes;
  }
};

__device__ __forceinline__ void put_pixel(uint8_t* d, uint32_t o) {
  d[0] = (uint8_t)(o & 0xffu); d[1] = (uint8_t)((o >> 8) & 0xffu); d[2] = (uint8_t)((o >> 16) & 0xffu);
}

// Tap fetch + blend of one pixel from its 1/32-px source coordinate; returns BGRx.
__device__ __forceinline__ uint32_t remap_one(const uint8_t* __restrict__ src, int W, int H, int ix, int iy, int ax, int ay,
                                              uint32_t border) {
  if ((unsigned)ix < (unsigned)(W - 3) && (unsigned)iy < (unsigned)(H - 1)) return blend_interior(src, W * 3, ix, iy, ax, ay);
  if (ix < -1 || ix >= W || iy < -1 || iy >= H) return border;
  uint8_t t3[3];
  remap_pixel(src, W, H, ix, iy, ax, ay, (int)(border & 0xffu), (int)((border >> 8) & 0xffu), (int)((border >> 16) & 0xffu), t3);
  return (uint32_t)t3[0] | ((uint32_t)t3[1] << 8) | ((uint32_t)t3[2] << 16);
}

// Owner of one pixel from its row's segment list.
__device__ __forceinline__ unsigned pixel_owner(const uint32_t* __restrict__ rs, int segcap, int px) {
  const uint4 e = __ldg(reinterpret_cast<const uint4*>(rs));
  const unsigned key = ((unsigned)px << 16) | 0xffffu;
  unsigned cur = e.x;
  if (e.y <= key) cur = e.y;                                 // sentinels never compare <= key
  if (e.z <= key) cur = e.z;
  if (e.w <= key) cur = e.w;
  return e.w != kSegSentinel ? seg_owner(rs, segcap, px) : (cur & 0xffffu);
}

// One pixel that only failed a GROUP condition (the group straddles two cells, footprints not
// adjacent, taps near the frame border): the float32 coordinate of the pixel's own cell is still exact
// outside the rounding band, only the tap fetch is per pixel.  Returns false -- nothing written -- when
// the pixel needs the float64 sequence after all (inside the band, cell without float32 form, irregular
// segment).
__device__ __forceinline__ bool medium_pixel(int px, int py, unsigned id, const uint8_t* __restrict__ src, uint8_t* dst,
                                             const CellFast* __restrict__ ffast, int W, int H, uint32_t border) {
  if (id == kSegIrregular) return false;
  if (id == kSegNone) {                                      // default map (W+1, H+1): border colour
    put_pixel(dst, border);
    return true;
  }
  const float4* cp = reinterpret_cast<const float4*>(ffast + id);
  const float4 q0 = __ldg(cp), q1 = __ldg(cp + 1), q2 = __ldg(cp + 2);
  const int4 q3 = __ldg(reinterpret_cast<const int4*>(cp + 3));
  int sx, sy;
  if (!medium_coords(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, __int_as_float(q3.w), __float_as_int(q2.z),
                     __float_as_int(q2.w), q3.x, q3.y, 0u, px, py, W, H, sx, sy))
    return false;
  put_pixel(dst, remap_one(src, W, H, sx >> 5, sy >> 5, sx & 31, sy & 31, border));
  return true;
}

// One pixel the exact way: owner from the segment list (or the per-pixel search for irregular segments),
// the reference's float64 remap sequence, the general tap fetch.
__device__ __noinline__ void slow_pixel(int px, int py, int f, const uint8_t* __restrict__ src, uint8_t* dst,
                                        const WarpTables T, int W, int H, uint32_t border) {
  const int tx = px / kTileW;
  const uint32_t* rs = T.rowseg + (((size_t)f * H + py) * T.tiles_x + tx) * T.segcap;
  const Cell* fcells = T.cells + (size_t)f * T.ncell;
  const unsigned id = pixel_owner(rs, T.segcap, px);
  float mx = (float)(W + 1), my = (float)(H + 1);           // mfs.py:983-984
  if (id == kSegIrregular) {
    const size_t tile = (size_t)f * T.tiles_x * T.tiles_y + (size_t)(py / kTileH) * T.tiles_x + tx;
    const int nraw = __ldg(T.tile_count + tile) & kCountMask;
    const bool overflow = nraw > kTileCap;
    const float2 m = resolve_pixel(fcells, overflow ? nullptr : T.tile_list + tile * kTileCap, overflow ? T.ncell : nraw,
                                   px, py, mx, my);
    mx = m.x; my = m.y;
  } else if (id != kSegNone) {
    const double2* hs = reinterpret_cast<const double2*>(fcells[id].Hsu);
    const double2 h01 = __ldg(hs), h23 = __ldg(hs + 1), h45 = __ldg(hs + 2), h67 = __ldg(hs + 3);
    const double y = (double)py;
    map_row(h01.x, h23.x, h23.y, h45.y, h67.x, (double)px, MF_MUL(y, h01.y), MF_MUL(y, h45.x), MF_MUL(y, h67.y), mx, my);
  }
  int ix, iy, ax, ay;
  remap_coords(mx, my, ix, iy, ax, ay);
  put_pixel(dst, remap_one(src, W, H, ix, iy, ax, ay, border));
}

// (B, G) pairs and R pair of pixel J of a group from the phase-0 words s0..s3 of one source row:
// bytes (3J, 3J+3), (3J+1, 3J+4) and (3J+2, 3J+5) of the row segment.
template <int J>
__device__ __forceinline__ void tap_pairs(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t& bg, uint32_t& r) {
  if (J == 0) { bg = __byte_perm(s0, s1, 0x4130); r = __byte_perm(s0, s1, 0x0052); }
  if (J == 1) { bg = __byte_perm(s0, s1, 0x7463); r = __byte_perm(s1, s2, 0x0041); }
  if (J == 2) { bg = __byte_perm(s1, s2, 0x6352); r = __byte_perm(s2, s2, 0x0030); }
  if (J == 3) { bg = __byte_perm(s2, s3, 0x5241); r = __byte_perm(s2, s3, 0x0063); }
}

// cv2.remap's blend of pixel J: (sum w_rc p_rc + 512) >> 10 with w_rc = wy_r * wx_c
template <int J>
__device__ __forceinline__ void blend_group_pixel(const uint32_t (&st)[4], const uint32_t (&sb)[4], unsigned ax, unsigned ay,
                                                  uint32_t& vb, uint32_t& vg, uint32_t& vr) {
  uint32_t tbg, tr, bbg, br;
  tap_pairs<J>(st[0], st[1], st[2], st[3], tbg, tr);
  tap_pairs<J>(sb[0], sb[1], sb[2], sb[3], bbg, br);
  const uint32_t wxp = ax * 0xffffu + 32u;                  // (32 - ax) | ax << 16
  const uint32_t wb = wxp * ay, wa = (wxp << 5) - wb;       // rows: ay, 32 - ay
  vb = __dp2a_lo(wb, bbg, __dp2a_lo(wa, tbg, 512u)) >> 10;
  vg = __dp2a_hi(wb, bbg, __dp2a_hi(wa, tbg, 512u)) >> 10;
  vr = __dp2a_lo(wb, br, __dp2a_lo(wa, tr, 512u)) >> 10;
}

// Float32 remap coordinates of a group of four adjacent pixels through one cell: nu / nv = rint(U), rint(V) +
// kRoundMagicBits as raw bits, du / dv = distance of U, V from that integer.  Returns true when all four pixels
// are outside the rounding band (NaN -- a cell without float32 form -- compares false).
__device__ __forceinline__ bool group_coords(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7,
                                             float a8, float thr_u, float thr_v, int cbx0, int cby0, int px0, int py,
                                             unsigned (&nu)[kPix], unsigned (&nv)[kPix], float (&du)[kPix], float (&dv)[kPix]) {
  const float fy = (float)(py - cby0);
  const float bx = fmaf(a1, fy, a2), by = fmaf(a4, fy, a5), bw = fmaf(a7, fy, a8);
  const float fx = (float)(px0 - cbx0);
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    const float x = fx + (float)j;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(a6, x, bw)));
    const float U = __fmul_rn(fmaf(a0, x, bx), r), V = __fmul_rn(fmaf(a3, x, by), r);
    const float tU = __fadd_rn(U, kRoundMagic), tV = __fadd_rn(V, kRoundMagic);
    du[j] = __fsub_rn(U, __fsub_rn(tU, kRoundMagic));
    dv[j] = __fsub_rn(V, __fsub_rn(tV, kRoundMagic));
    nu[j] = __float_as_uint(tU); nv[j] = __float_as_uint(tV);
  }
  const float wu = fmaxf(fmaxf(fabsf(du[0]), fabsf(du[1])), fmaxf(fabsf(du[2]), fabsf(du[3])));
  const float wv = fmaxf(fmaxf(fabsf(dv[0]), fabsf(dv[1])), fmaxf(fabsf(dv[2]), fabsf(dv[3])));
  // fmaxf drops a NaN operand, but a cell either has finite coefficients (no NaN at all) or none (four NaNs)
  return wu <= thr_u && wv <= thr_v;
}

// List entry of one group of four pixels: lane (5 bits) | row within the warp's rows (5 bits) | pixel mask (4 bits).
// Two lists per warp: groups with pixels for the per-pixel tap fetch, groups with pixels for the float64 path.
static constexpr int kSlowList = 192;                       // pixels gathered for one float64 batch (>= 128 + 32 + 32)

template <int kRows>
struct WarpScratch {
  uint16_t med[32 * kRows];                                  // every group of a warp's kRows rows fits
  uint16_t ex[32 * kRows + 32];                              // + one late entry per lane (see warp_rows)
  uint16_t slow[kSlowList];
};

// The rows [y_first, y_first + nrows) x the 128 columns [X0, X0 + 128) of frame f, restricted to the columns
// [x_lo, x_hi] the caller needs (multiples of four are NOT required; groups that do not touch the range are
// skipped).  kShared: `sink` is the CTA's shared-memory tile (sink_shared = its shared-window address).
template <bool kShared>
__device__ __forceinline__ void warp_rows(const uint8_t* __restrict__ src, const WarpTables& T, int f, int W, int H, int X0,
                                          int y_first, int nrows, int x_lo, int x_hi, const PixelSink& sink,
                                          unsigned sink_shared, uint32_t border, uint16_t* __restrict__ q_med,
                                          uint16_t* __restrict__ q_ex, uint16_t* __restrict__ q_slow, int lane) {
  const int px0 = X0 + lane * kPix;
  int npx = min(kPix, W - px0);                              // <= 0: this lane has no pixels
  if (px0 > x_hi || px0 + kPix - 1 < x_lo) npx = 0;
  const unsigned fcell0 = (unsigned)f * (unsigned)T.ncell;   // < 2^32: at most 65535 frames x 65533 cells
  const unsigned pitch = (unsigned)W * 3u;
  const bool word_store = kShared || (((sink.pitch & 3u) == 0u) && ((reinterpret_cast<uintptr_t>(sink.at(px0, y_first)) & 3u) == 0u));
  // running pointers: one add per row instead of a 64-bit multiply chain
  // Every lane loads an owner every row, lanes without pixels from a clamped (valid) entry: a lane-dependent branch
  // around the load would make the other lanes' path rewrite the load's destination register and stall on it.
  const size_t own_stride = (size_t)T.tiles_x * 32;
  const uint16_t* own = T.lane_owner + ((size_t)f * H + y_first) * own_stride + min(px0 >> 2, T.tiles_x * 32 - 1);
  uint8_t* drow = kShared ? nullptr : sink.at(px0, y_first);
  // the shared-memory tile holds BGRx pixels: a group is one aligned 16-byte store, and the resize phase reads taps as
  // whole words (no byte phases)
  unsigned srow = kShared ? sink_shared + (unsigned)(y_first - sink.oy) * sink.pitch + (unsigned)(px0 - sink.ox) * 4u : 0u;
  const CellFast* ffast = T.fast + fcell0;
  const unsigned lt_mask = (1u << lane) - 1u;
  // the lists' shared-window addresses, once: generic pointers would be re-derived (S2R + LEA) every row
  const unsigned med_sh = (unsigned)__cvta_generic_to_shared(q_med), ex_sh = (unsigned)__cvta_generic_to_shared(q_ex);
  int n_med = 0, n_ex = 0;                                   // entries in the warp's two lists (warp-uniform)

  // the owner of the next row's group is requested one row ahead (it comes from L2)
  unsigned own_next = kSegNone;
  if (nrows > 0) own_next = __ldg(own);
#pragma unroll 1
  for (int r = 0; r < nrows; ++r, own += own_stride, drow += kShared ? 0 : sink.pitch, srow += kShared ? sink.pitch : 0u) {
    const int py = y_first + r;
    const unsigned id = own_next;
    if (r + 1 < nrows) own_next = __ldg(own + own_stride);      // warp-uniform condition
    unsigned flag = 0u;                                      // bits 0-3: per-pixel tap fetch, bits 4-7: float64 path
    bool fast_group = false;
    unsigned nu[kPix], nv[kPix];
    int ix0 = 0, iy0 = 0;
    if (npx > 0) {
      if (id >= kSegStraddle || npx < kPix) {                // one test on the hot path; the rare owners sort themselves out here
        const unsigned all = (1u << npx) - 1u;
        if (id == kSegIrregular) {
          flag = all << 4;
        } else if (id != kSegNone || npx < kPix) {
          flag = all;
        } else {                                             // no cell: map (W+1, H+1), border colour
          if (kShared) {
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(srow), "r"(border));
          } else {
            uint32_t o[kPix] = {border, border, border, border};
            store_bgr4(drow, o, kPix, word_store);
          }
        }
      } else {
        // the cell's parameters come from L1 every row: cheaper than keeping 16 registers alive across the gather
        const float4* cp = reinterpret_cast<const float4*>(ffast + id);
        const float4 q0 = __ldg(cp), q1 = __ldg(cp + 1), q2 = __ldg(cp + 2);
        const int4 q3 = __ldg(reinterpret_cast<const int4*>(cp + 3));
        const float thr_u = q2.y, thr_v = __int_as_float(q3.w);
        float du[kPix], dv[kPix];
        const bool all_safe = group_coords(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, thr_u, thr_v,
                                           __float_as_int(q2.z), __float_as_int(q2.w), px0, py, nu, nv, du, dv);
        const unsigned bu = nu[0] & ~31u, bv = nv[0] & ~31u;
        const unsigned spread = (nu[1] - bu - 32u) | (nu[2] - bu - 64u) | (nu[3] - bu - 96u) | (nv[1] - bv) | (nv[2] - bv) |
                                (nv[3] - bv);
        ix0 = (int)(bu - kRoundMagicBits + (unsigned)q3.x) >> 5;
        iy0 = (int)(bv - kRoundMagicBits + (unsigned)q3.y) >> 5;
        // adjacent footprints (pixel j reads source columns ix0+j, ix0+j+1 of rows iy0, iy0+1) with every tap -- and
        // the 5-word row reads -- inside the frame
        fast_group = spread < 32u && (unsigned)ix0 <= (unsigned)(W - 8) && (unsigned)iy0 <= (unsigned)(H - 2);
        if (!all_safe) {                                     // rare: which of the four are inside the rounding band
          unsigned bad = 0u;
#pragma unroll
          for (int j = 0; j < kPix; ++j)
            if (!(fabsf(du[j]) <= thr_u && fabsf(dv[j]) <= thr_v)) bad |= 1u << j;
          flag = bad << 4;
          if (!(thr_u >= 0.0f)) { flag = 15u << 4; fast_group = false; }     // cell without float32 form
        }
        if (!fast_group) flag |= 15u & ~(flag >> 4);         // only the group shape failed: per-pixel tap fetch
      }
    }
    if (fast_group) {
      const unsigned bu = nu[0] & ~31u, bv = nv[0] & ~31u;
      const uintptr_t p0 = reinterpret_cast<uintptr_t>(src) + (unsigned)iy0 * pitch + (unsigned)ix0 * 3u;
      const uintptr_t p1 = p0 + pitch;
      const unsigned f0 = (unsigned)(p0 & 3u), f1 = (unsigned)(p1 & 3u);
      const uint32_t* w0 = reinterpret_cast<const uint32_t*>(p0 & ~(uintptr_t)3);
      const uint32_t* w1 = reinterpret_cast<const uint32_t*>(p1 & ~(uintptr_t)3);
      uint32_t t[5], b[5];
#ifdef MF_EXP_TAP_MASK
      // EXPERIMENT (results invalid): the taps of the fast path read from shared memory with 32-bit addresses, as if
      // the tile's source box had been staged there for free (TMA): an upper bound on what staging can bring.
      {
        const unsigned o0 = ((unsigned)iy0 * pitch + (unsigned)ix0 * 3u);
        const unsigned s0 = o0 & (unsigned)MF_EXP_TAP_MASK & ~3u, s1 = (o0 + pitch) & (unsigned)MF_EXP_TAP_MASK & ~3u;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t[i]) : "r"(s0 + 4u * i));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b[i]) : "r"(s1 + 4u * i));
        }
      }
#else
#pragma unroll
      for (int i = 0; i < 4; ++i) { t[i] = __ldg(w0 + i); b[i] = __ldg(w1 + i); }
#ifndef MF_FAST_NO_PREFETCH
      // the next output row reads this window one source row further down: its bottom row is the only new line
      if (iy0 + 2 < H) asm volatile("prefetch.global.L1 [%0];" ::"l"(p1 + pitch + 8u));
#endif
      t[4] = f0 >= 2u ? __ldg(w0 + 4) : 0u;
      b[4] = f1 >= 2u ? __ldg(w1 + 4) : 0u;
#endif
      uint32_t st[4], sb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st[i] = __funnelshift_r(t[i], t[i + 1], f0 * 8u);
        sb[i] = __funnelshift_r(b[i], b[i + 1], f1 * 8u);
      }
      uint32_t vb[kPix], vg[kPix], vr[kPix];
      blend_group_pixel<0>(st, sb, nu[0] - bu, nv[0] - bv, vb[0], vg[0], vr[0]);
      blend_group_pixel<1>(st, sb, nu[1] - bu - 32u, nv[1] - bv, vb[1], vg[1], vr[1]);
      blend_group_pixel<2>(st, sb, nu[2] - bu - 64u, nv[2] - bv, vb[2], vg[2], vr[2]);
      blend_group_pixel<3>(st, sb, nu[3] - bu - 96u, nv[3] - bv, vb[3], vg[3], vr[3]);
      if (kShared) {
        uint32_t p[kPix];
#pragma unroll
        for (int j = 0; j < kPix; ++j) p[j] = __byte_perm(__byte_perm(vb[j], vg[j], 0x0040), vr[j], 0x0410);   // B G R x
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(srow), "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]));
      } else if (word_store) {
        const uint32_t o0 = __byte_perm(__byte_perm(vb[0], vg[0], 0x0040), __byte_perm(vr[0], vb[1], 0x0040), 0x5410);
        const uint32_t o1 = __byte_perm(__byte_perm(vg[1], vr[1], 0x0040), __byte_perm(vb[2], vg[2], 0x0040), 0x5410);
        const uint32_t o2 = __byte_perm(__byte_perm(vr[2], vb[3], 0x0040), __byte_perm(vg[3], vr[3], 0x0040), 0x5410);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(drow);
        __stcs(d32 + 0, o0); __stcs(d32 + 1, o1); __stcs(d32 + 2, o2);
      } else {
#pragma unroll
        for (int j = 0; j < kPix; ++j) { drow[3 * j] = (uint8_t)vb[j]; drow[3 * j + 1] = (uint8_t)vg[j]; drow[3 * j + 2] = (uint8_t)vr[j]; }
      }
    }
    // one list entry per group that declined (or partly declined) the fast path: slots from ballots
    if (__any_sync(0xffffffffu, flag != 0u)) {
      const unsigned mm = flag & 15u, xm = flag >> 4;
      const unsigned bal_m = __ballot_sync(0xffffffffu, mm != 0u), bal_x = __ballot_sync(0xffffffffu, xm != 0u);
      const unsigned tag = (unsigned)lane | ((unsigned)r << 5);
      if (mm != 0u)
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(med_sh + 2u * (unsigned)(n_med + __popc(bal_m & lt_mask))), "h"((unsigned short)(tag | (mm << 10))) : "memory");
      if (xm != 0u)
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(ex_sh + 2u * (unsigned)(n_ex + __popc(bal_x & lt_mask))), "h"((unsigned short)(tag | (xm << 10))) : "memory");
      n_med += __popc(bal_m);
      n_ex += __popc(bal_x);
    }
  }
  if ((n_med | n_ex) == 0) return;
  __syncwarp();
  // ---- what the fast path declined, handled by this warp right away (its source rows are still in L1 / L2) ----
  // (1) per-pixel tap fetch: eight groups x four pixels per step.  A pixel that turns out to be inside the rounding
  //     band after all (rare) is remembered by its lane and joins the float64 list as a one-pixel group at the end;
  //     a lane that already remembers one has the warp work the remembered pixels off first.
  const int sub = lane >> 2, j = lane & 3;
  unsigned pend = 0u;                                        // remembered entry (mask != 0) or 0
  for (int i0 = 0; i0 < n_med; i0 += 8) {
    const int i = i0 + sub;
    bool failed = false;
    unsigned e = 0u;
    if (i < n_med) {
      e = q_med[i];
      if ((e >> (10 + j)) & 1u) {
        const int qx = X0 + (int)(e & 31u) * kPix + j, qy = y_first + (int)((e >> 5) & 31u);
        const uint32_t* rsq = T.rowseg + (((size_t)f * H + qy) * T.tiles_x + qx / kTileW) * T.segcap;
        failed = !medium_pixel(qx, qy, pixel_owner(rsq, T.segcap, qx), src, sink.at(qx, qy), ffast, W, H, border);
      }
    }
    if (__any_sync(0xffffffffu, failed)) {
      if (__any_sync(0xffffffffu, failed && pend != 0u)) {
        if (pend != 0u) {
          const int qx = X0 + (int)(pend & 31u) * kPix + (__ffs((int)(pend >> 10)) - 1), qy = y_first + (int)((pend >> 5) & 31u);
          slow_pixel(qx, qy, f, src, sink.at(qx, qy), T, W, H, border);
        }
        pend = 0u;
      }
      if (failed) pend = (e & 1023u) | (1u << (10 + j));
    }
  }
  {
    const unsigned pb = __ballot_sync(0xffffffffu, pend != 0u);
    if (pend != 0u) q_ex[n_ex + __popc(pb & lt_mask)] = (uint16_t)pend;
    n_ex += __popc(pb);
  }
  __syncwarp();
  // (2) float64 path: the groups' flagged pixels are compacted into a pixel list, 32 groups at a time, and the list
  //     is worked off whenever it cannot take another batch
  int ns = 0;                                                // pixels waiting in q_slow (warp-uniform)
  for (int i0 = 0; i0 < n_ex; i0 += 32) {
    const int i = i0 + lane;
    const unsigned e = i < n_ex ? q_ex[i] : 0u;
    unsigned m = (e >> 10) & 15u;
    const int cnt = __popc(m);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int slot = ns + incl - cnt;
    while (m != 0u) {
      const int b = __ffs((int)m) - 1;
      m &= m - 1u;
      q_slow[slot++] = (uint16_t)(((e & 31u) * kPix + (unsigned)b) | (((e >> 5) & 31u) << 7));
    }
    ns += total;
    __syncwarp();
    if (ns > kSlowList - 128 || i0 + 32 >= n_ex) {
      for (int k = lane; k < ns; k += 32) {
        const unsigned p = q_slow[k];
        const int qx = X0 + (int)(p & 127u), qy = y_first + (int)(p >> 7);
        slow_pixel(qx, qy, f, src, sink.at(qx, qy), T, W, H, border);
      }
      ns = 0;
      __syncwarp();
    }
  }
}

#ifndef MF_FAST_MINBLOCKS
#define MF_FAST_MINBLOCKS 4
#endif
// Stabilized frames to global memory.  CTA = 128 x kFastTileH output pixels, warp = 128 x kFastRows.
__global__ void __launch_bounds__(kWarpThreads, MF_FAST_MINBLOCKS) warp_fast_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out, const WarpTables T, int W, int H, uint32_t border) {
  __shared__ WarpScratch<kFastRows> scratch[kWarpThreads / 32];
  const int f = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y_first = blockIdx.y * kFastTileH + warp * kFastRows;
  const int nrows = min(kFastRows, H - y_first);
  if (nrows <= 0) return;
  PixelSink sink;
  sink.base = frames_out + (size_t)f * H * W * 3;
  sink.pitch = (unsigned)W * 3u; sink.ox = 0; sink.oy = 0; sink.px_bytes = 3u;
  warp_rows<false>(frames_in + (size_t)f * H * W * 3, T, f, W, H, blockIdx.x * kTileW, y_first, nrows, 0, W - 1, sink, 0u,
                   border, scratch[warp].med, scratch[warp].ex, scratch[warp].slow, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// Fused pass B: stabilized tile in shared memory -> cv2.resize of the crop window (mfs.py:1111-1157)
// ---------------------------------------------------------------------------------------------------------------
#ifndef MF_FUSED_TILE_H
#define MF_FUSED_TILE_H 64
#endif
static constexpr int kOutTileW = 120;                       // final-frame pixels per CTA: 30 groups of four
static constexpr int kOutTileH = MF_FUSED_TILE_H;           // eight rows per warp in the resize phase
static_assert(kOutTileH % 8 == 0 && kOutTileH >= 8 && kOutTileH <= 120, "resize phase: whole rows per warp");
static constexpr int kStabRows = kOutTileH + 1;             // scale <= 1: at most one source row more than output rows
static constexpr int kStabRowsPerWarp = (kStabRows + 7) / 8;
static constexpr int kStabPitch = kTileW * 4;               // 512 B: 128 stabilized BGRx pixels per row
static_assert(kStabRowsPerWarp <= 16 && kStabRowsPerWarp * 8 >= kStabRows, "rows of the stabilized tile split over eight warps");

struct FusedShared {
  uint8_t tile[kStabRows * kStabPitch + 32];               // + slack: a zero-weight tap may read past the last pixel
  WarpScratch<kStabRowsPerWarp> scratch[kWarpThreads / 32];
#ifdef MF_EXP_EXTRA_SMEM
  uint8_t exp_box[MF_EXP_EXTRA_SMEM];                        // EXPERIMENT: the shared memory a staged source box would take
#endif
};

// Horizontal pass of cv2.resize for the four pixels of a thread on one row of the shared-memory tile (BGRx words):
// (a0 * p[c0] + a1 * p[c0 + 1]) >> 4 per channel.  woff[j] = byte offset of pixel j's left tap in the row.
__device__ __forceinline__ void fused_hsum(unsigned row_shared, const unsigned (&woff)[kPix], const uint32_t (&wx)[kPix],
                                           RowSums& out) {
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    uint32_t t0, t1;
    const unsigned a = row_shared + woff[j];
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t0) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(t1) : "r"(a));
    const uint32_t bg = __byte_perm(t0, t1, 0x5140), rr = __byte_perm(t0, t1, 0x0062);   // B0 B1 G0 G1 | R0 R1 . .
    out.v[j][0] = __dp2a_lo(wx[j], bg, 0u) >> 4;
    out.v[j][1] = __dp2a_hi(wx[j], bg, 0u) >> 4;
    out.v[j][2] = __dp2a_lo(wx[j], rr, 0u) >> 4;
  }
}

#ifndef MF_FUSED_MINBLOCKS
#define MF_FUSED_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kWarpThreads, MF_FUSED_MINBLOCKS) warp_fused_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out, const WarpTables T, int table_frame0, int W, int H,
    uint32_t border, const int32_t* __restrict__ enc4, const int4* __restrict__ xtab, const int4* __restrict__ ytab) {
  extern __shared__ __align__(16) unsigned char fused_raw[];
  FusedShared& sm = *reinterpret_cast<FusedShared*>(fused_raw);
  int left, top, right, bottom;
  if (!decode_crop(enc4, W, H, left, top, right, bottom)) return;      // host raises once it reads the rectangle back
  const int fl = blockIdx.z;                                 // frame within this call
  const int f = table_frame0 + fl;                           // frame within the tables
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ox0 = blockIdx.x * kOutTileW, oy0 = blockIdx.y * kOutTileH;
  const int ox1 = min(W, ox0 + kOutTileW) - 1, oy1 = min(H, oy0 + kOutTileH) - 1;
  // stabilized pixels this tile's taps touch (crop-relative tap indices from the resize tables)
  const int xa = left + __ldg(xtab + ox0).x, xb = min(right, left + __ldg(xtab + ox1).x + 1);
  const int ya = top + __ldg(ytab + oy0).x, yb = top + __ldg(ytab + oy1).y;
  const int X0 = xa & ~3;
  const uint8_t* src = frames_in + (size_t)fl * H * W * 3;
  const unsigned tile_shared = (unsigned)__cvta_generic_to_shared(sm.tile);
  // the tile's stabilized rows split evenly over the eight warps (7 or 8 each for a typical 61-64 rows): the warps
  // meet at a barrier, so the longest share is what counts.  (Per-warp ready flags instead of the barrier -- a warp
  // only needs the rows of two or three warps -- were measured: 3.73 ms vs 3.60 ms per 300 frames, the polling costs
  // more issue slots than the barrier wastes.)
  const int total = yb - ya + 1, base = total >> 3, extra = total & 7;
  {
    PixelSink sink;
    sink.base = sm.tile; sink.pitch = kStabPitch; sink.ox = X0; sink.oy = ya; sink.px_bytes = 4u;
    const int y_first = ya + warp * base + min(warp, extra);
    const int nrows = base + (warp < extra ? 1 : 0);
#ifdef MF_EXP_NO_WARP                  // timing experiment only: cost of the resize phase alone
    if (false)
#else
    if (nrows > 0)
#endif
      warp_rows<true>(src, T, f, W, H, X0, y_first, nrows, xa, xb, sink, tile_shared, border, sm.scratch[warp].med,
                      sm.scratch[warp].ex, sm.scratch[warp].slow, lane);
  }
  __syncthreads();
#ifdef MF_EXP_NO_RESIZE                // timing experiment only: cost of the stabilization phase alone
  if (sm.tile[threadIdx.x] != 7 || true) return;
#endif
  // ---- resize phase: warp = 8 output rows x 120 pixels, four adjacent output pixels per thread ----
  const int px0 = ox0 + lane * kPix;
  if (lane * kPix >= kOutTileW || px0 > ox1) return;
  const int y_out0 = oy0 + warp * (kOutTileH / 8);
  const int y_out1 = min(oy1, y_out0 + kOutTileH / 8 - 1);
  if (y_out0 > y_out1) return;
  const int npx = min(kPix, ox1 - px0 + 1);
  const unsigned out_pitch = (unsigned)W * 3u;
  uint8_t* drow = frames_out + ((size_t)fl * H * W + (size_t)y_out0 * W + px0) * 3;
  uint32_t wx[kPix];
  unsigned woff[kPix];
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    const int4 xt = __ldg(xtab + min(px0 + j, W - 1));
    // the right tap of a pixel clamped at the crop's last column carries weight 0: reading c0 + 1 is harmless
    wx[j] = (uint32_t)xt.z | ((uint32_t)xt.w << 16);
    woff[j] = (unsigned)(left + xt.x - X0) * 4u;
  }
  const bool word_store = npx == kPix && (out_pitch & 3u) == 0u && (reinterpret_cast<uintptr_t>(drow) & 3u) == 0u;
  RowSums P, Q;
  int rp = -1, rq = -1;
  bool swapped = false;                                      // false: P is the top row, Q the bottom row
  auto vertical = [&](const RowSums& Tt, const RowSums& M, const int4& yt) {
    const uint32_t b0s = (uint32_t)yt.z << 16, b1s = (uint32_t)yt.w << 16;
    uint32_t v[kPix][3];
#pragma unroll
    for (int jj = 0; jj < kPix; ++jj)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)   // a0 + a1 <= 2049 and b0 + b1 <= 2049 bound the result by 255: no clamp needed
        v[jj][ch] = (__umulhi(b0s, Tt.v[jj][ch]) + __umulhi(b1s, M.v[jj][ch]) + 2u) >> 2;
    if (word_store) {
      uint32_t* d32 = reinterpret_cast<uint32_t*>(drow);
      __stcs(d32 + 0, __byte_perm(__byte_perm(v[0][0], v[0][1], 0x0040), __byte_perm(v[0][2], v[1][0], 0x0040), 0x5410));
      __stcs(d32 + 1, __byte_perm(__byte_perm(v[1][1], v[1][2], 0x0040), __byte_perm(v[2][0], v[2][1], 0x0040), 0x5410));
      __stcs(d32 + 2, __byte_perm(__byte_perm(v[2][2], v[3][0], 0x0040), __byte_perm(v[3][1], v[3][2], 0x0040), 0x5410));
    } else {
#pragma unroll
      for (int jj = 0; jj < kPix; ++jj)
        if (jj < npx) { drow[3 * jj] = (uint8_t)v[jj][0]; drow[3 * jj + 1] = (uint8_t)v[jj][1]; drow[3 * jj + 2] = (uint8_t)v[jj][2]; }
    }
  };
  // Tt holds source row rt in the top role, M holds rm in the bottom role; returns true when the roles swapped
  auto step = [&](RowSums& Tt, RowSums& M, int& rt, int& rm, int r0, int r1, const int4& yt) -> bool {
    if (r0 != rt && r0 == rm && r1 != rm) {                  // the usual move: old bottom becomes top, one new row
      fused_hsum(tile_shared + (unsigned)(r1 - ya) * kStabPitch, woff, wx, Tt);
      rt = r1;
      vertical(M, Tt, yt);
      return true;
    }
    if (r0 != rt) {
      if (r0 == rm) Tt = M; else fused_hsum(tile_shared + (unsigned)(r0 - ya) * kStabPitch, woff, wx, Tt);
      rt = r0;
    }
    if (r1 != rm) {
      if (r1 == rt) M = Tt; else fused_hsum(tile_shared + (unsigned)(r1 - ya) * kStabPitch, woff, wx, M);
      rm = r1;
    }
    vertical(Tt, M, yt);
    return false;
  };
#ifndef MF_FUSED_NO_PROLOGUE
  {
    // prologue: the first output row's top source row enters in the bottom role, so that the first step is already
    // "the usual move" (one new row) and the general form stays out of the common path
    rq = top + __ldg(ytab + y_out0).x;
    fused_hsum(tile_shared + (unsigned)(rq - ya) * kStabPitch, woff, wx, Q);
  }
#endif
  int4 yt_next = __ldg(ytab + y_out0);
  for (int py = y_out0; py <= y_out1; ++py, drow += out_pitch) {
    const int4 yt = yt_next;
    if (py < y_out1) yt_next = __ldg(ytab + py + 1);         // requested a row ahead (warp-uniform condition)
    const int r0 = top + yt.x, r1 = top + yt.y;              // warp-uniform
    const bool flip = swapped ? step(Q, P, rq, rp, r0, r1, yt) : step(P, Q, rp, rq, r0, r1, yt);
    swapped = swapped != flip;
  }
}

}  // namespace mf
