// Fast path of the mesh warp (included by warp.cu).
//
// The generic kernel (warp_kernel) decides, for every group of four pixels, which cell owns it by
// screening candidates, and evaluates the reference's float64 remap sequence for every pixel: about
// 190 instructions per pixel, issue-bound at 10 % of the HBM roofline.  The fast path moves every
// decision that does not depend on pixel data out of the pixel loop:
//
//   cell_spans_kernel   : one thread per (frame, cell, four rows of the cell's box): exact member interval of
//                         the cell on a row (span_of_row: four half-planes + the exact test for a pixel
//                         within rounding noise of a boundary).
//   row_segments_kernel : one thread per (frame, row, 128-px tile): the candidates' intervals resolved by
//                         "the last cell written wins" into <= 16 sorted segments (first x, cell id), and
//                         the owner of each lane's group of four pixels (what the pixel kernel reads).
//   warp_fast_kernel    : CTA = 128 x 120 output pixels, warp = 128 x 15, four adjacent pixels per thread
//                         and row.  Per row a thread reads its group's owner (one 16-bit load, requested a
//                         row ahead), the cell's 64-byte parameter block from L1, evaluates the remap
//                         coordinates in float32 in box-centred form (error < eps, see mf_math.cuh) and
//                         keeps rint() only when the value is outside the rounding band; when the four
//                         footprints are adjacent the two source rows are read once as 4-5 aligned words
//                         each (the next row's new line is prefetched into L1), brought to phase 0 with
//                         funnel shifts, and the taps are regrouped with constant-selector PRMTs for the
//                         dp2a blend.
//                         Everything else -- pixels inside the rounding band, groups that straddle a
//                         segment, non-adjacent footprints, taps outside the frame, crop-edge
//                         candidates, irregular cells -- is listed in the warp's shared-memory queue and
//                         handled by the same warp right after its rows, lanes packed: medium_pixel()
//                         (float32 coordinate, per-pixel tap fetch) for pixels that only failed a group
//                         condition, slow_pixel() (the reference's float64 sequence) for the rest.
//
// Results are identical to warp_kernel (tests/test_gpu_parity.py compares both with the oracle).
#pragma once

namespace mf {

// Member interval of every cell on every row of its support box.  span_tab[(f * ncell + id) * span_rows + (y - by0)] = a | b << 16 (a > b: no member pixel;
// kSpanIrregular: the row cannot be described by an interval with certainty).  Cells whose box is taller than
// span_rows are treated as irregular by row_segments_kernel.
static constexpr uint32_t kSpanIrregular = 0xffffffffu;
static constexpr uint32_t kSpanEmpty = 0x00000001u;          // a = 1, b = 0

static constexpr int kSpanRowsPerThread = 4;                 // independent rows per thread: four chains in flight

__global__ void __launch_bounds__(128) cell_spans_kernel(const Cell* __restrict__ cells, const CellSpan* __restrict__ spans,
                                                         int64_t ncells_total, int span_rows, uint32_t* __restrict__ span_tab) {
  const int chunks = (span_rows + kSpanRowsPerThread - 1) / kSpanRowsPerThread;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ncells_total * chunks) return;
  const int64_t cid = idx / chunks;
  const int rl0 = (int)(idx - cid * chunks) * kSpanRowsPerThread;
  const Cell& c = cells[cid];
  const int4 box = __ldg(reinterpret_cast<const int4*>(&c));
  if (box.x > box.z || box.y + rl0 > box.w) return;
  const CellSpan& sp = spans[cid];
  const bool regular = sp.regular != 0;
  uint32_t* out = span_tab + cid * span_rows;
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

__global__ void __launch_bounds__(128) row_segments_kernel(
    const Cell* __restrict__ cells, const uint32_t* __restrict__ span_tab, int span_rows, const int* __restrict__ tile_count,
    const uint16_t* __restrict__ tile_list, int nf, int W, int H, int ncell, int tiles_x, int tiles_y, int segcap,
    uint32_t* __restrict__ rowseg, uint4* __restrict__ lane_owner) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)nf * H * tiles_x;
  if (idx >= total) return;
  const int tx = (int)(idx % tiles_x);
  const int64_t fy = idx / tiles_x;
  const int y = (int)(fy % H), f = (int)(fy / H);
  const int x0 = tx * kTileW, x1 = min(W - 1, x0 + kTileW - 1);
  const size_t tile = (size_t)f * tiles_x * tiles_y + (size_t)(y / kTileH) * tiles_x + tx;
  const int nraw = __ldg(tile_count + tile) & kCountMask;
  uint32_t* out = rowseg + (size_t)idx * segcap;
  SegBuilder sb;
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
      if (a <= b) sb.cover(a, b, (unsigned)id[j], segcap);
      if (sb.overflow) { irregular = true; continue; }
      if (sb.done()) done = true;
    }
  }
  int ns = irregular ? -1 : sb.finish(segcap);
  if (ns < 0) {
    out[0] = ((unsigned)x0 << 16) | kSegIrregular;
    for (int i = 1; i < segcap; ++i) out[i] = kSegSentinel;
  } else {
    for (int i = 0; i < segcap; ++i) out[i] = i < ns ? sb.seg[i] : kSegSentinel;
  }
  // owner of each lane's group of four pixels (what the pixel kernel reads: one 16-bit load per lane and row)
  uint32_t packed[16];
  if (ns < 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) packed[i] = kSegIrregular | (kSegIrregular << 16);
  } else {
    int si = 0;
    unsigned cur = sb.seg[0] & 0xffffu;
    int next_x = ns > 1 ? (int)(sb.seg[1] >> 16) : 0x7fffffff;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const int g0 = x0 + kPix * l;
      while (next_x <= g0) {                                 // segments are sorted: si only moves forward
        ++si;
        cur = sb.seg[si] & 0xffffu;
        next_x = si + 1 < ns ? (int)(sb.seg[si + 1] >> 16) : 0x7fffffff;
      }
      const unsigned o = next_x <= g0 + kPix - 1 ? kSegStraddle : cur;
      if (l & 1) packed[l >> 1] |= o << 16; else packed[l >> 1] = o;
    }
  }
  uint4* lo = lane_owner + (size_t)idx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) lo[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
}

// Tap fetch + blend + 3-byte store of one pixel from its 1/32-px source coordinate.
__device__ __forceinline__ void remap_store_pixel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst_frame, int px,
                                                  int py, int W, int H, int ix, int iy, int ax, int ay, uint32_t border) {
  uint32_t o;
  if ((unsigned)ix < (unsigned)(W - 3) && (unsigned)iy < (unsigned)(H - 1)) {
    o = blend_interior(src, W * 3, ix, iy, ax, ay);
  } else if (ix < -1 || ix >= W || iy < -1 || iy >= H) {
    o = border;
  } else {
    uint8_t t3[3];
    remap_pixel(src, W, H, ix, iy, ax, ay, (int)(border & 0xffu), (int)((border >> 8) & 0xffu), (int)((border >> 16) & 0xffu), t3);
    o = (uint32_t)t3[0] | ((uint32_t)t3[1] << 8) | ((uint32_t)t3[2] << 16);
  }
  uint8_t* d = dst_frame + ((size_t)py * W + px) * 3;
  d[0] = (uint8_t)(o & 0xffu); d[1] = (uint8_t)((o >> 8) & 0xffu); d[2] = (uint8_t)((o >> 16) & 0xffu);
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
// the pixel needs the float64 sequence after all (inside the band, crop-edge candidate, cell without
// float32 form, irregular segment).
__device__ __forceinline__ bool medium_pixel(int px, int py, unsigned id, const uint8_t* __restrict__ src,
                                             uint8_t* __restrict__ dst_frame, const CellFast* __restrict__ ffast, int W,
                                             int H, uint32_t border) {
  if (id == kSegIrregular) return false;
  if (id == kSegNone) {                                      // default map (W+1, H+1): border colour, no crop hit
    if (dst_frame != nullptr) {
      uint8_t* d = dst_frame + ((size_t)py * W + px) * 3;
      d[0] = (uint8_t)(border & 0xffu); d[1] = (uint8_t)((border >> 8) & 0xffu); d[2] = (uint8_t)((border >> 16) & 0xffu);
    }
    return true;
  }
  const float4* cp = reinterpret_cast<const float4*>(ffast + id);
  const float4 q0 = __ldg(cp), q1 = __ldg(cp + 1), q2 = __ldg(cp + 2);
  const int4 q3 = __ldg(reinterpret_cast<const int4*>(cp + 3));
  int sx, sy;
  if (!medium_coords(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, __int_as_float(q3.w), __float_as_int(q2.z),
                     __float_as_int(q2.w), q3.x, q3.y, (unsigned)q3.z, px, py, W, H, sx, sy))
    return false;
  if (dst_frame != nullptr) remap_store_pixel(src, dst_frame, px, py, W, H, sx >> 5, sy >> 5, sx & 31, sy & 31, border);
  return true;
}

// One pixel the exact way: owner from the segment list (or the per-pixel search for irregular segments),
// the reference's float64 remap sequence, the four crop-edge searches, the general tap fetch.
__device__ __noinline__ void slow_pixel(int px, int py, int f, const uint8_t* __restrict__ src,
                                        uint8_t* __restrict__ dst_frame, const Cell* __restrict__ fcells,
                                        const int* __restrict__ tile_count,
                                        const uint16_t* __restrict__ tile_list, const uint32_t* __restrict__ rowseg,
                                        int segcap, int32_t* __restrict__ crop_out, int W, int H, int ncell, int tiles_x,
                                        int tiles_y, uint32_t border) {
  const int tx = px / kTileW;
  const uint32_t* rs = rowseg + (((size_t)f * H + py) * tiles_x + tx) * segcap;
  const unsigned id = pixel_owner(rs, segcap, px);
  float mx = (float)(W + 1), my = (float)(H + 1);           // mfs.py:983-984
  if (id == kSegIrregular) {
    const size_t tile = (size_t)f * tiles_x * tiles_y + (size_t)(py / kTileH) * tiles_x + tx;
    const int nraw = __ldg(tile_count + tile) & kCountMask;
    const bool overflow = nraw > kTileCap;
    const float2 m = resolve_pixel(fcells, overflow ? nullptr : tile_list + tile * kTileCap, overflow ? ncell : nraw,
                                   px, py, mx, my);
    mx = m.x; my = m.y;
  } else if (id != kSegNone) {
    const double2* hs = reinterpret_cast<const double2*>(fcells[id].Hsu);
    const double2 h01 = __ldg(hs), h23 = __ldg(hs + 1), h45 = __ldg(hs + 2), h67 = __ldg(hs + 3);
    const double y = (double)py;
    map_row(h01.x, h23.x, h23.y, h45.y, h67.x, (double)px, MF_MUL(y, h01.y), MF_MUL(y, h45.x), MF_MUL(y, h67.y), mx, my);
  }
  int32_t* cr = crop_out + 4 * f;                           // mfs.py:1075-1098; plain read first: most hits do not improve
  if (mx > -1.0f && mx < 1.0f && px > cr[0]) atomicMax(cr + 0, px);
  if (my > -1.0f && my < 1.0f && py > cr[1]) atomicMax(cr + 1, py);
  if (mx > (float)(W - 2) && mx < (float)W && px < cr[2]) atomicMin(cr + 2, px);
  if (my > (float)(H - 2) && my < (float)H && py < cr[3]) atomicMin(cr + 3, py);
  if (dst_frame == nullptr) return;
  int ix, iy, ax, ay;
  remap_coords(mx, my, ix, iy, ax, ay);
  remap_store_pixel(src, dst_frame, px, py, W, H, ix, iy, ax, ay, border);
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

#ifndef MF_FAST_MINBLOCKS
#define MF_FAST_MINBLOCKS 4
#endif
template <bool kBoundsOnly>
__global__ void __launch_bounds__(kWarpThreads, MF_FAST_MINBLOCKS) warp_fast_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out, const Cell* __restrict__ cells,
    const CellFast* __restrict__ fast, const int* __restrict__ tile_count, const uint16_t* __restrict__ tile_list,
    const uint32_t* __restrict__ rowseg, const uint16_t* __restrict__ lane_owner, int segcap,
    int32_t* __restrict__ crop_out, int W, int H, int ncell, int tiles_x, int tiles_y, uint32_t border) {
  __shared__ uint16_t warp_queue[kWarpThreads / 32][kTileW * kFastRows];   // every pixel of a warp fits
  const int f = blockIdx.z, tx = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int px0 = tx * kTileW + lane * kPix;
  const int y_first = blockIdx.y * kFastTileH + warp * kFastRows;
  const int npx = min(kPix, W - px0);                        // <= 0: this lane has no pixels
  const uint8_t* src = frames_in + (size_t)f * H * W * 3;
  uint8_t* dstf = kBoundsOnly ? nullptr : frames_out + (size_t)f * H * W * 3;
  const unsigned fcell0 = (unsigned)f * (unsigned)ncell;     // < 2^32: at most 65535 frames x 65533 cells
  const unsigned pitch = (unsigned)W * 3u;
  const bool word_store = ((pitch & 3u) == 0u) && ((reinterpret_cast<uintptr_t>(dstf) & 3u) == 0u);
  // running pointers: one add per row instead of a 64-bit multiply chain
  const size_t own_stride = (size_t)tiles_x * 32;
  const uint16_t* own = lane_owner + (((size_t)f * H + y_first) * tiles_x + tx) * 32 + lane;
  uint8_t* drow = kBoundsOnly ? nullptr : dstf + ((size_t)y_first * W + px0) * 3;

  // four bits per row, newest row in the low bits: pixel j of the row takes the float64 path / the per-pixel tap fetch
  unsigned long long exmask = 0ull, medmask = 0ull;
  int rows_done = 0;

  // the owner of the next row's group is requested one row ahead (it comes from L2)
  unsigned own_next = kSegNone;
  if (y_first < H && npx > 0) own_next = __ldg(own);
#pragma unroll 1
  for (int r = 0; r < kFastRows; ++r, own += own_stride, drow += pitch) {
    const int py = y_first + r;
    if (py >= H) break;
    ++rows_done;
    exmask <<= 4; medmask <<= 4;
    if (npx <= 0) continue;
    const unsigned id = own_next;
    if (r + 1 < kFastRows && py + 1 < H) own_next = __ldg(own + own_stride);
    if (kBoundsOnly) {                                       // only tiles that hold a border cell can produce a hit
      const size_t tile = (size_t)f * tiles_x * tiles_y + (size_t)(py / kTileH) * tiles_x + tx;
      if (!(__ldg(tile_count + tile) & kEdgeFlag)) continue;
    }
    unsigned push = 0u, med = 0u;                            // bit j: pixel j -> float64 path / per-pixel tap fetch
    bool fast_group = false;
    unsigned nu[kPix], nv[kPix];
    int ix0 = 0, iy0 = 0, base_x = 0, base_y = 0;
    if (id >= kSegStraddle || npx < kPix) {                  // one test on the hot path; the rare owners sort themselves out here
      if (id == kSegIrregular) {
        push = (1u << npx) - 1u;
      } else if (id != kSegNone || npx < kPix) {
        med = (1u << npx) - 1u;
      } else if (!kBoundsOnly) {                             // no cell: map (W+1, H+1), border colour, no crop hit
        uint32_t o[kPix] = {border, border, border, border};
        store_bgr4(drow, o, kPix, word_store);
      }
    } else {
      // the cell's parameters come from L1 every row: cheaper than keeping 16 registers alive across the gather
      const float4* cp = reinterpret_cast<const float4*>(fast + (size_t)(fcell0 + id));
      const float4 q0 = __ldg(cp), q1 = __ldg(cp + 1), q2 = __ldg(cp + 2);
      const int4 q3 = __ldg(reinterpret_cast<const int4*>(cp + 3));
      const float a0 = q0.x, a1 = q0.y, a2 = q0.z, a3 = q0.w, a4 = q1.x, a5 = q1.y, a6 = q1.z, a7 = q1.w, a8 = q2.x, thr = q2.y;
      const int cbx0 = __float_as_int(q2.z), cby0 = __float_as_int(q2.w);
      base_x = q3.x; base_y = q3.y;
      const unsigned flags = (unsigned)q3.z;
      const float thr_v = __int_as_float(q3.w);
      // no early exit for cells without float32 form (thr < 0: their coefficients are zero, every pixel comes out
      // "in the band"): a branch here would serialise the four parameter loads behind the first one
      const unsigned bad = fast_group_coords(a0, a1, a2, a3, a4, a5, a6, a7, a8, thr, thr_v, cbx0, cby0, px0, py, nu, nv);
      bool edge;
      push = fast_group_plan(nu, nv, bad, base_x, base_y, flags, W, H, kBoundsOnly, ix0, iy0, fast_group, edge);
      if (push == 15u && !edge) { push = bad; med = 15u & ~bad; }   // only the group shape failed
      if (thr < 0.0f) { push = 15u; med = 0u; fast_group = false; }
    }
    exmask |= push;
    medmask |= med;
    if (!kBoundsOnly && fast_group) {
      const unsigned bu = nu[0] & ~31u, bv = nv[0] & ~31u;
      const uintptr_t p0 = reinterpret_cast<uintptr_t>(src) + (unsigned)iy0 * pitch + (unsigned)ix0 * 3u;
      const uintptr_t p1 = p0 + pitch;
      const unsigned f0 = (unsigned)(p0 & 3u), f1 = (unsigned)(p1 & 3u);
      const uint32_t* w0 = reinterpret_cast<const uint32_t*>(p0 & ~(uintptr_t)3);
      const uint32_t* w1 = reinterpret_cast<const uint32_t*>(p1 & ~(uintptr_t)3);
      uint32_t t[5], b[5];
#pragma unroll
      for (int i = 0; i < 4; ++i) { t[i] = __ldg(w0 + i); b[i] = __ldg(w1 + i); }
#ifndef MF_FAST_NO_PREFETCH
      // the next output row reads this window one source row further down: its bottom row is the only new line
      if (iy0 + 2 < H) asm volatile("prefetch.global.L1 [%0];" ::"l"(p1 + pitch + 8u));
#endif
      t[4] = f0 >= 2u ? __ldg(w0 + 4) : 0u;
      b[4] = f1 >= 2u ? __ldg(w1 + 4) : 0u;
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
      if (word_store) {
        uint32_t* d32 = reinterpret_cast<uint32_t*>(drow);
        __stcs(d32 + 0, __byte_perm(__byte_perm(vb[0], vg[0], 0x0040), __byte_perm(vr[0], vb[1], 0x0040), 0x5410));
        __stcs(d32 + 1, __byte_perm(__byte_perm(vg[1], vr[1], 0x0040), __byte_perm(vb[2], vg[2], 0x0040), 0x5410));
        __stcs(d32 + 2, __byte_perm(__byte_perm(vr[2], vb[3], 0x0040), __byte_perm(vg[3], vr[3], 0x0040), 0x5410));
      } else {
#pragma unroll
        for (int j = 0; j < kPix; ++j) { drow[3 * j] = (uint8_t)vb[j]; drow[3 * j + 1] = (uint8_t)vg[j]; drow[3 * j + 2] = (uint8_t)vr[j]; }
      }
    }
  }
  // ---- what the fast path declined, handled by this warp right away (its source rows are still in L1 / L2):
  //      one scan for both masks (counts packed as 16-bit halves), entries in the warp's shared-memory list ----
  const int mine = __popcll(medmask) | (__popcll(exmask) << 16);
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int totals = __shfl_sync(0xffffffffu, incl, 31);
  if (totals == 0) return;
  const int n_med = totals & 0xffff, n_ex = totals >> 16;
  uint16_t* wq = warp_queue[warp];                          // [0, n_med): per-pixel tap fetch; [n_med, n_med + n_ex): float64
  {
    int slot = (incl & 0xffff) - (mine & 0xffff);
    while (medmask != 0ull) {
      const int b = __ffsll((long long)medmask) - 1;
      medmask &= medmask - 1ull;
      wq[slot++] = (uint16_t)((lane * kPix + (b & 3)) | ((rows_done - 1 - (b >> 2)) << 7));
    }
    slot = n_med + (incl >> 16) - (mine >> 16);
    while (exmask != 0ull) {
      const int b = __ffsll((long long)exmask) - 1;
      exmask &= exmask - 1ull;
      wq[slot++] = (uint16_t)((lane * kPix + (b & 3)) | ((rows_done - 1 - (b >> 2)) << 7));
    }
  }
  __syncwarp();
  const CellFast* ffast2 = fast + (size_t)f * ncell;
  int n_failed = 0;                                          // pixels whose tap fetch declined, compacted to the list's front
  for (int i0 = 0; i0 < n_med; i0 += 32) {
    const int i = i0 + lane;
    bool failed = false;
    unsigned e = 0u;
    if (i < n_med) {
      e = wq[i];
      const int qx = tx * kTileW + (int)(e & 127u), qy = y_first + (int)(e >> 7);
      const uint32_t* rsq = rowseg + (((size_t)f * H + qy) * tiles_x + tx) * segcap;
      failed = !medium_pixel(qx, qy, pixel_owner(rsq, segcap, qx), src, dstf, ffast2, W, H, border);
    }
    const unsigned fb = __ballot_sync(0xffffffffu, failed);  // rare: the pixel is in the rounding band after all
    __syncwarp();                                            // every lane has read its entry: the front may be reused
    if (failed) wq[n_failed + __popc(fb & ((1u << lane) - 1u))] = (uint16_t)e;
    n_failed += __popc(fb);
  }
  __syncwarp();
  const Cell* fcells = cells + (size_t)f * ncell;
  for (int i = lane; i < n_failed + n_ex; i += 32) {
    const unsigned e = wq[i < n_failed ? i : n_med + (i - n_failed)];
    slow_pixel(tx * kTileW + (int)(e & 127u), y_first + (int)(e >> 7), f, src, dstf, fcells, tile_count, tile_list, rowseg,
               segcap, crop_out, W, H, ncell, tiles_x, tiles_y, border);
  }
}

}  // namespace mf
