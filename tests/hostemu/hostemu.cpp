// TEST INFRASTRUCTURE ONLY -- never loaded by the meshflow_b200 package.
//
// Drives the __host__ __device__ per-element functions of meshflow_b200/csrc/mf_math.cuh on the CPU
// (g++ -ffp-contract=off, so no FMA contraction, like the explicit _rn intrinsics on the device) so
// that the CPU-only test-suite can compare the very arithmetic the kernels run against the oracle.
// It contains no kernels and no parallel plumbing; the GPU tests cover those.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cmath>
#include <utility>
#include "../../meshflow_b200/csrc/mf_math.cuh"

extern "C" {

// per-feature row windows and column ranges: left/right are [n][R+1] (left > right when empty)
void emu_feature_ranges(const double* fx, const double* fy, int n, int W, int H, int R, int C, int er, int ec,
                        int* top, int* bot, int* left, int* right) {
  for (int i = 0; i < n; ++i) {
    mf::FeatureCell fc = mf::feature_cell(fx[i], fy[i], W, H, R, C, er);
    top[i] = fc.top; bot[i] = fc.bot;
    for (int vr = 0; vr <= R; ++vr) {
      int l = 1, r = 0;
      if (vr >= fc.top && vr <= fc.bot) mf::col_range(fc, vr, C, er, ec, l, r);
      left[i * (R + 1) + vr] = l; right[i * (R + 1) + vr] = r;
    }
  }
}

void emu_persp(const double* M, const double* x, const double* y, int n, double* ox, double* oy) {
  for (int i = 0; i < n; ++i) mf::persp(M, x[i], y[i], ox[i], oy[i]);
}

void emu_key_roundtrip(const double* v, int n, uint64_t* keys, double* back) {
  for (int i = 0; i < n; ++i) { keys[i] = mf::key_of(v[i]); back[i] = mf::value_of(keys[i]); }
}

float emu_median9(const float* v) { float t[9]; memcpy(t, v, sizeof(t)); return mf::median9(t); }

double emu_lambda(const double* Hm, int W, int H, int definition) { return mf::adaptive_lambda(Hm, W, H, definition); }

int emu_sizeof_cell() { return (int)sizeof(mf::Cell); }

// cells of one frame: rest_xy [V,2] float32, delta = s - u [V,2] float64
void emu_cell_setup(const float* vertex_xy, const double* u, const double* s, int W, int H, int R, int C,
                    mf::Cell* cells) {
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      const int vidx[4] = {r * (C + 1) + c, r * (C + 1) + c + 1, (r + 1) * (C + 1) + c, (r + 1) * (C + 1) + c + 1};
      double rest[8], stab[8];
      for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 2; ++d) {
          const int v = vidx[k];
          const double rv = (double)vertex_xy[2 * v + d];
          rest[2 * k + d] = rv;
          stab[2 * k + d] = (double)(float)(rv + (s[2 * v + d] - u[2 * v + d]));
        }
      mf::cell_setup(rest, stab, W, H, cells[r * C + c]);
    }
}

// the warp kernel's per-pixel logic with every cell as candidate (descending id, box prune)
void emu_warp_frame(const uint8_t* src, const mf::Cell* cells, int ncell, int W, int H, int bb, int bg, int br,
                    uint8_t* dst, float* map_xy, int* crop4, int use_box) {
  int e_left = 0, e_top = 0, e_right = W - 1, e_bottom = H - 1;
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      float mx = (float)(W + 1), my = (float)(H + 1);
      for (int id = ncell - 1; id >= 0; --id) {
        const mf::Cell& c = cells[id];
        const bool box = px >= c.bx0 && px <= c.bx1 && py >= c.by0 && py <= c.by1;
        if (!(box || !use_box)) continue;
        // the kernel's order of decisions: float32 screen first, float64 test only when it abstains
        const float fy = (float)(py - c.by0);
        int sc = (use_box && c.feps >= 0.0f)
                     ? mf::cell_screen(c, (float)(px - c.bx0), fmaf(c.fm[1], fy, c.fm[2]), fmaf(c.fm[4], fy, c.fm[5]),
                                       fmaf(c.fm[7], fy, c.fm[8]))
                     : -1;
        if (sc < 0) sc = mf::cell_inside(c, (double)px, (double)py) ? 1 : 0;
        if (sc == 1) {
          const double y = (double)py;
          mf::cell_map_row(c, (double)px, y * c.Hsu[1], y * c.Hsu[4], y * c.Hsu[7], mx, my);
          break;
        }
      }
      if (map_xy) { map_xy[2 * ((size_t)py * W + px)] = mx; map_xy[2 * ((size_t)py * W + px) + 1] = my; }
      if (mx > -1.0f && mx < 1.0f && px > e_left) e_left = px;
      if (mx > (float)(W - 2) && mx < (float)W && px < e_right) e_right = px;
      if (my > -1.0f && my < 1.0f && py > e_top) e_top = py;
      if (my > (float)(H - 2) && my < (float)H && py < e_bottom) e_bottom = py;
      int ix, iy, ax, ay;
      mf::remap_coords(mx, my, ix, iy, ax, ay);
      mf::remap_pixel(src, W, H, ix, iy, ax, ay, bb, bg, br, dst + ((size_t)py * W + px) * 3);
    }
  crop4[0] = e_left; crop4[1] = e_top; crop4[2] = e_right; crop4[3] = e_bottom;
}

// Screening audit over one frame: out[0] = (pixel, cell) pairs examined (pixel inside the cell's box),
// out[1] = pairs the float32 screen left to the float64 test, out[2] = pairs where a screen decision
// contradicts the float64 test (must be 0), out[3] = cells that may be screened at all.
void emu_screen_audit(const mf::Cell* cells, int ncell, int W, int H, long long* out) {
  out[0] = out[1] = out[2] = out[3] = 0;
  for (int id = 0; id < ncell; ++id) {
    const mf::Cell& c = cells[id];
    if (c.feps >= 0.0f) out[3]++;
    if (c.bx0 > c.bx1) continue;
    for (int py = c.by0; py <= c.by1; ++py) {
      const float fy = (float)(py - c.by0);
      const float bx = fmaf(c.fm[1], fy, c.fm[2]), by = fmaf(c.fm[4], fy, c.fm[5]), bw = fmaf(c.fm[7], fy, c.fm[8]);
      for (int px = c.bx0; px <= c.bx1; ++px) {
        const int sc = mf::cell_screen(c, (float)(px - c.bx0), bx, by, bw);
        const bool exact = mf::cell_inside(c, (double)px, (double)py);
        out[0]++;
        if (sc < 0) out[1]++;
        else if ((sc == 1) != exact) out[2]++;
      }
    }
  }
}

// NOTE: the kernel's crop search starts from "no hit"; a hit at column 0 and no hit are the same value.
void emu_crop_resize(const uint8_t* src, int W, int H, int left, int top, int right, int bottom, uint8_t* dst) {
  const int sw = right - left + 1, sh = bottom - top + 1;
  std::vector<int> x0(W), x1(W), a0(W), a1(W);
  for (int x = 0; x < W; ++x) mf::resize_coef(x, sw, W, true, x0[x], x1[x], a0[x], a1[x]);
  for (int y = 0; y < H; ++y) {
    int r0, r1, b0, b1;
    mf::resize_coef(y, sh, H, false, r0, r1, b0, b1);
    const uint8_t* p0 = src + (size_t)(top + r0) * W * 3;
    const uint8_t* p1 = src + (size_t)(top + r1) * W * 3;
    for (int x = 0; x < W; ++x)
      for (int ch = 0; ch < 3; ++ch)
        dst[((size_t)y * W + x) * 3 + ch] = (uint8_t)mf::resize_blend(
            p0[(left + x0[x]) * 3 + ch], p0[(left + x1[x]) * 3 + ch], p1[(left + x0[x]) * 3 + ch],
            p1[(left + x1[x]) * 3 + ch], a0[x], a1[x], b0, b1);
  }
}


// ---- fast path of the warp (meshflow_b200/csrc/warp_fast.cuh), same decisions on the CPU -----------
static void emu_cells_all(const float* vertex_xy, const double* u, const double* s, int W, int H, int R, int C,
                          std::vector<mf::Cell>& cells, std::vector<mf::CellFast>& fast, std::vector<mf::CellSpan>& spans) {
  cells.resize(R * C); fast.resize(R * C); spans.resize(R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      const int vidx[4] = {r * (C + 1) + c, r * (C + 1) + c + 1, (r + 1) * (C + 1) + c, (r + 1) * (C + 1) + c + 1};
      double rest[8], stab[8];
      for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 2; ++d) {
          const int v = vidx[k];
          const double rv = (double)vertex_xy[2 * v + d];
          rest[2 * k + d] = rv;
          stab[2 * k + d] = (double)(float)(rv + (s[2 * v + d] - u[2 * v + d]));
        }
      const int id = r * C + c;
      mf::cell_setup(rest, stab, W, H, cells[id]);
      mf::cell_fast_setup(cells[id], (int)floor(fmin(rest[0], rest[4])), (int)ceil(fmax(rest[2], rest[6])),
                          (int)floor(fmin(rest[1], rest[3])), (int)ceil(fmax(rest[5], rest[7])), W, H, fast[id], spans[id]);
      if (fast[id].thr_u >= 0.0f) cells[id].edge_flags |= mf::kMapMonotone;      // as cell_setup_kernel does
    }
}

// Every (cell, row) span against the exact membership test, pixel by pixel over the cell's box.
// out[0] = rows examined, out[1] = rows span_of_row declined (status 2), out[2] = pixels where the span
// disagrees with cell_inside (must be 0), out[3] = regular cells.
void emu_span_audit(const float* vertex_xy, const double* u, const double* s, int W, int H, int R, int C, long long* out) {
  std::vector<mf::Cell> cells; std::vector<mf::CellFast> fast; std::vector<mf::CellSpan> spans;
  emu_cells_all(vertex_xy, u, s, W, H, R, C, cells, fast, spans);
  out[0] = out[1] = out[2] = out[3] = 0;
  for (int id = 0; id < R * C; ++id) {
    const mf::Cell& c = cells[id];
    if (!spans[id].regular) continue;
    out[3]++;
    for (int y = c.by0; y <= c.by1; ++y) {
      int a = 1, b = 0;
      const int st = mf::span_of_row(c, spans[id], y, c.bx0, c.bx1, a, b);
      out[0]++;
      if (st == 2) { out[1]++; continue; }
      if (st == 1) { a = 1; b = 0; }
      for (int x = c.bx0; x <= c.bx1; ++x)
        if (mf::cell_inside(c, (double)x, (double)y) != (x >= a && x <= b)) out[2]++;
    }
    // rows outside the box hold no member pixel (the box is a superset of the mask)
  }
}

// stats: [0] pixels through the shared-window gather, [1] slow pixels, [2] pixels in the rounding band,
// [3] irregular row segments, [4] row segments, [5] groups sent whole to the slow path
void emu_warp_frame_fast(const uint8_t* src, const float* vertex_xy, const double* u, const double* s, int W, int H,
                         int R, int C, int bb, int bg, int br, uint8_t* dst, int* crop4, int bounds_only, long long* stats) {
  const int kTileW = 128, kTileH = 8, kTileCap = 48;
  std::vector<mf::Cell> cells; std::vector<mf::CellFast> fast; std::vector<mf::CellSpan> spans;
  emu_cells_all(vertex_xy, u, s, W, H, R, C, cells, fast, spans);
  const int ncell = R * C, tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const int segcap = (W / C >= 48) ? 8 : mf::kSegMax;
  std::vector<std::vector<int>> lists(tiles_x * tiles_y);
  for (int id = ncell - 1; id >= 0; --id) {                 // descending id
    const mf::Cell& c = cells[id];
    if (c.bx0 > c.bx1) continue;
    for (int ty = c.by0 / kTileH; ty <= c.by1 / kTileH; ++ty)
      for (int tx = c.bx0 / kTileW; tx <= c.bx1 / kTileW; ++tx) lists[ty * tiles_x + tx].push_back(id);
  }
  std::vector<unsigned> rowseg((size_t)H * tiles_x * segcap);
  for (int i = 0; i < 6; ++i) stats[i] = 0;
  for (int y = 0; y < H; ++y)
    for (int tx = 0; tx < tiles_x; ++tx) {
      const std::vector<int>& l = lists[(y / kTileH) * tiles_x + tx];
      const int x0 = tx * kTileW, x1 = std::min(W - 1, x0 + kTileW - 1);
      unsigned* out = &rowseg[((size_t)y * tiles_x + tx) * segcap];
      stats[4]++;
      auto build = [&](auto& sb) {
        sb.begin(x0, x1);
        bool irregular = (int)l.size() > kTileCap;
        for (size_t k = 0; k < l.size() && !irregular; ++k) {
          const int id = l[k];
          const mf::Cell& c = cells[id];
          if (y < c.by0 || y > c.by1 || x1 < c.bx0 || x0 > c.bx1) continue;
          if (!spans[id].regular) { irregular = true; break; }
          int a, b;
          const int st = mf::span_of_row(c, spans[id], y, std::max(x0, c.bx0), std::min(x1, c.bx1), a, b);
          if (st == 2) { irregular = true; break; }
          if (st == 0) sb.cover(a, b, (unsigned)id);
          if (sb.overflow) { irregular = true; break; }
          if (sb.done()) break;
        }
        const int ns = irregular ? -1 : sb.finish();
        if (ns < 0) { out[0] = ((unsigned)x0 << 16) | mf::kSegIrregular; for (int i = 1; i < segcap; ++i) out[i] = mf::kSegSentinel; stats[3]++; }
        else for (int i = 0; i < segcap; ++i) out[i] = sb.seg[i];
      };
      if (segcap == 8) { mf::SegBuilder<8> sb; build(sb); } else { mf::SegBuilder<16> sb; build(sb); }
    }
  // crop edges from the row segments of tiles that hold a border cell (crop_edges_kernel)
  int crop[4] = {0, 0, W - 1, H - 1};
  for (int y = 0; y < H; ++y)
    for (int tx = 0; tx < tiles_x; ++tx) {
      const std::vector<int>& l = lists[(y / kTileH) * tiles_x + tx];
      bool edge_tile = false;
      for (int id : l) edge_tile = edge_tile || (cells[id].edge_flags & mf::kEdgeAny) != 0u;
      if (!edge_tile) continue;
      const int x0 = tx * kTileW, x1 = std::min(W - 1, x0 + kTileW - 1);
      const unsigned* rs = &rowseg[((size_t)y * tiles_x + tx) * segcap];
      if ((rs[0] & 0xffffu) == mf::kSegIrregular) {
        const bool overflow = (int)l.size() > kTileCap;
        const int n = overflow ? ncell : (int)l.size();
        for (int px = x0; px <= x1; ++px) {
          float mx = (float)(W + 1), my = (float)(H + 1);
          for (int k = 0; k < n; ++k) {
            const mf::Cell& c = cells[overflow ? ncell - 1 - k : l[k]];
            if (px < c.bx0 || px > c.bx1 || y < c.by0 || y > c.by1) continue;
            if (mf::cell_inside(c, (double)px, (double)y)) { mf::cell_map(c, (double)px, (double)y, mx, my); break; }
          }
          if (mx > -1.0f && mx < 1.0f && px > crop[0]) crop[0] = px;
          if (my > -1.0f && my < 1.0f && y > crop[1]) crop[1] = y;
          if (mx > (float)(W - 2) && mx < (float)W && px < crop[2]) crop[2] = px;
          if (my > (float)(H - 2) && my < (float)H && y < crop[3]) crop[3] = y;
        }
        continue;
      }
      int ns = 0;
      while (ns < segcap && rs[ns] != mf::kSegSentinel) ++ns;
      for (int i = 0; i < ns; ++i) {
        const unsigned id = rs[i] & 0xffffu;
        if (id == mf::kSegNone) continue;
        const int xa = (int)(rs[i] >> 16), xb = i + 1 < ns ? (int)(rs[i + 1] >> 16) - 1 : x1;
        mf::segment_crop_edges(cells[id], xa, xb, y, W, H, crop);
      }
    }
  for (int i = 0; i < 4; ++i) crop4[i] = crop[i];
  if (bounds_only) return;
  std::vector<std::pair<int, int>> queue, medium;           // float64 path / per-pixel tap fetch
  const int bord[3] = {bb, bg, br};
  for (int py = 0; py < H; ++py)
    for (int px0 = 0; px0 < W; px0 += 4) {
      const int npx = std::min(4, W - px0), tx = px0 / kTileW;
      const unsigned* rs = &rowseg[((size_t)py * tiles_x + tx) * segcap];
      bool strad;
      const unsigned id = mf::seg_group_owner(rs, segcap, px0, strad);
      unsigned push = 0u, med = 0u; bool fg = false; unsigned nu[4], nv[4]; int ix0 = 0, iy0 = 0;
      if (id == mf::kSegIrregular) push = (1u << npx) - 1u;
      else if (strad || npx < 4) med = (1u << npx) - 1u;
      else if (id == mf::kSegNone) {
        if (!bounds_only) for (int j = 0; j < 4; ++j) for (int ch = 0; ch < 3; ++ch) dst[((size_t)py * W + px0 + j) * 3 + ch] = (uint8_t)bord[ch];
      } else {
        const mf::CellFast& cf = fast[id];
        if (cf.thr_u < 0.0f) push = 15u;
        else {
          const unsigned bad = mf::fast_group_coords(cf.a[0], cf.a[1], cf.a[2], cf.a[3], cf.a[4], cf.a[5], cf.a[6], cf.a[7],
                                                     cf.a[8], cf.thr_u, cf.thr_v, cf.bx0, cf.by0, px0, py, nu, nv);
          stats[2] += __builtin_popcount(bad);
          bool edge; push = mf::fast_group_plan(nu, nv, bad, cf.base_x, cf.base_y, 0u, W, H, false, ix0, iy0, fg, edge);
          if (push == 15u) { push = bad; med = 15u & ~bad; }
        }
      }
      if (push == 15u || (npx < 4 && push)) stats[5]++;
      for (int j = 0; j < 4; ++j) if (push & (1u << j)) queue.push_back({px0 + j, py});
      for (int j = 0; j < 4; ++j) if (med & (1u << j)) medium.push_back({px0 + j, py});
      if (fg && !bounds_only) {
        const unsigned bu = nu[0] & ~31u, bv = nv[0] & ~31u;
        for (int j = 0; j < 4; ++j) {
          const int ax = (int)(nu[j] - bu) - 32 * j, ay = (int)(nv[j] - bv);
          const uint8_t* r0 = src + ((size_t)iy0 * W + ix0 + j) * 3;
          const uint8_t* r1 = r0 + (size_t)W * 3;
          for (int ch = 0; ch < 3; ++ch)
            dst[((size_t)py * W + px0 + j) * 3 + ch] = (uint8_t)mf::blend4(r0[ch], r0[3 + ch], r1[ch], r1[3 + ch], ax, ay);
        }
        stats[0] += 4;
      }
    }
  for (auto& q : medium) {                                  // medium_pixel of the kernel
    const int px = q.first, py = q.second, tx = px / kTileW;
    const unsigned id = mf::seg_owner(&rowseg[((size_t)py * tiles_x + tx) * segcap], segcap, px);
    if (id == mf::kSegIrregular) { queue.push_back(q); continue; }
    if (id == mf::kSegNone) {
      if (!bounds_only) for (int ch = 0; ch < 3; ++ch) dst[((size_t)py * W + px) * 3 + ch] = (uint8_t)bord[ch];
      continue;
    }
    const mf::CellFast& cf = fast[id];
    int sx, sy;
    if (!mf::medium_coords(cf.a[0], cf.a[1], cf.a[2], cf.a[3], cf.a[4], cf.a[5], cf.a[6], cf.a[7], cf.a[8], cf.thr_u, cf.thr_v,
                           cf.bx0, cf.by0, cf.base_x, cf.base_y, 0u, px, py, W, H, sx, sy)) { queue.push_back(q); continue; }
    if (!bounds_only) mf::remap_pixel(src, W, H, sx >> 5, sy >> 5, sx & 31, sy & 31, bb, bg, br, dst + ((size_t)py * W + px) * 3);
  }
  stats[1] = (long long)queue.size();
  for (auto& q : queue) {
    const int px = q.first, py = q.second, tx = px / kTileW;
    const unsigned id = mf::seg_owner(&rowseg[((size_t)py * tiles_x + tx) * segcap], segcap, px);
    float mx = (float)(W + 1), my = (float)(H + 1);
    if (id == mf::kSegIrregular) {
      const std::vector<int>& l = lists[(py / kTileH) * tiles_x + tx];
      const bool overflow = (int)l.size() > kTileCap;
      const int n = overflow ? ncell : (int)l.size();
      for (int k = 0; k < n; ++k) {
        const mf::Cell& c = cells[overflow ? ncell - 1 - k : l[k]];
        if (px < c.bx0 || px > c.bx1 || py < c.by0 || py > c.by1) continue;
        if (mf::cell_inside(c, (double)px, (double)py)) { mf::cell_map(c, (double)px, (double)py, mx, my); break; }
      }
    } else if (id != mf::kSegNone) {
      mf::cell_map(cells[id], (double)px, (double)py, mx, my);
    }
    int ix, iy, ax, ay;
    mf::remap_coords(mx, my, ix, iy, ax, ay);
    mf::remap_pixel(src, W, H, ix, iy, ax, ay, bb, bg, br, dst + ((size_t)py * W + px) * 3);
  }
}


// per cell: out[6*id..] = regular, bounded, thr_u, thr_v, box width, box height
void emu_cell_fast_info(const float* vertex_xy, const double* u, const double* s, int W, int H, int R, int C, double* out) {
  std::vector<mf::Cell> cells; std::vector<mf::CellFast> fast; std::vector<mf::CellSpan> spans;
  emu_cells_all(vertex_xy, u, s, W, H, R, C, cells, fast, spans);
  for (int id = 0; id < R * C; ++id) {
    out[6 * id] = spans[id].regular; out[6 * id + 1] = cells[id].bounded; out[6 * id + 2] = fast[id].thr_u;
    out[6 * id + 3] = fast[id].thr_v; out[6 * id + 4] = cells[id].bx1 - cells[id].bx0 + 1; out[6 * id + 5] = cells[id].by1 - cells[id].by0 + 1;
  }
}


// SegBuilder on its own: candidates (descending priority order as given) with intervals [a, b] on the tile
// [x0, x1]; writes the sorted segments, returns their number (-1: more than cap) and per-pixel owners from
// seg_owner / seg_group_owner for comparison with a brute-force resolution.
int emu_resolve_segments(int x0, int x1, int n, const int* a, const int* b, const int* ids, int cap, unsigned* seg_out,
                         unsigned* owner_px, unsigned* owner_group) {
  unsigned seg[mf::kSegMax];
  auto build = [&](auto& sb) -> int {
    sb.begin(x0, x1);
    for (int k = 0; k < n && !sb.overflow && !sb.done(); ++k) {
      const int lo = std::max(a[k], x0), hi = std::min(b[k], x1);
      if (lo <= hi) sb.cover(lo, hi, (unsigned)ids[k]);
    }
    const int ns = sb.finish();
    for (int i = 0; i < mf::kSegMax; ++i) seg[i] = (ns >= 0 && i < cap) ? sb.seg[i < cap ? i : 0] : mf::kSegSentinel;
    return ns;
  };
  int ns;
  if (cap == 8) { mf::SegBuilder<8> sb; ns = build(sb); } else { mf::SegBuilder<16> sb; ns = build(sb); }
  if (ns < 0) return -1;
  for (int i = 0; i < mf::kSegMax; ++i) seg_out[i] = seg[i];
  for (int x = x0; x <= x1; ++x) owner_px[x - x0] = mf::seg_owner(seg, cap, x);
  for (int g = x0; g <= x1; g += 4) {
    bool strad;
    const unsigned o = mf::seg_group_owner(seg, cap, g, strad);
    owner_group[(g - x0) / 4] = strad ? mf::kSegStraddle : o;
  }
  return ns;
}


// Audit of the float32 remap coordinates over every member pixel of every cell with a float32 form:
// out[0] = pixels examined, out[1] = pixels inside the rounding band (float64 path), out[2] = SAFE pixels whose
// rint() differs from the reference's 1/32-px coordinate (must be 0); ratio[0] = largest observed
// |float32 value - float64 value| over the cell's eps (how conservative the bound is; must stay < 1).
void emu_fast_coord_audit(const float* vertex_xy, const double* u, const double* s, int W, int H, int R, int C,
                          long long* out, double* ratio) {
  std::vector<mf::Cell> cells; std::vector<mf::CellFast> fast; std::vector<mf::CellSpan> spans;
  emu_cells_all(vertex_xy, u, s, W, H, R, C, cells, fast, spans);
  out[0] = out[1] = out[2] = 0; ratio[0] = 0.0;
  for (int id = 0; id < R * C; ++id) {
    const mf::Cell& c = cells[id];
    const mf::CellFast& cf = fast[id];
    if (cf.thr_u < 0.0f || c.bx0 > c.bx1) continue;
    for (int py = c.by0; py <= c.by1; ++py)
      for (int px = c.bx0; px <= c.bx1; ++px) {
        if (!mf::cell_inside(c, (double)px, (double)py)) continue;
        float mx, my;
        mf::cell_map(c, (double)px, (double)py, mx, my);
        int ix, iy, ax, ay;
        mf::remap_coords(mx, my, ix, iy, ax, ay);
        const int sx_ref = ix * 32 + ax, sy_ref = iy * 32 + ay;
        const float fy = (float)(py - cf.by0), fx = (float)(px - cf.bx0);
        unsigned nu, nv;
        const bool safe = mf::fast_coords(cf.a[0], cf.a[3], cf.a[6], fmaf(cf.a[1], fy, cf.a[2]), fmaf(cf.a[4], fy, cf.a[5]),
                                          fmaf(cf.a[7], fy, cf.a[8]), fx, cf.thr_u, cf.thr_v, nu, nv);
        out[0]++;
        if (!safe) { out[1]++; continue; }
        const int sx = (int)(nu - mf::kRoundMagicBits) + cf.base_x, sy = (int)(nv - mf::kRoundMagicBits) + cf.base_y;
        if (sx != sx_ref || sy != sy_ref) out[2]++;
        // observed error of the float32 value against the float64 map (in 1/32-px units), relative to eps = 0.5 - thr - tie
        const float Uf = fmaf(cf.a[0], fx, fmaf(cf.a[1], fy, cf.a[2])) / fmaf(cf.a[6], fx, fmaf(cf.a[7], fy, cf.a[8]));
        const double w = (double)px * c.Hsu[6] + (double)py * c.Hsu[7] + 1.0;
        const double ue = 32.0 * (((double)px * c.Hsu[0] + (double)py * c.Hsu[1] + c.Hsu[2]) / w) - (double)cf.base_x;
        const double band = 0.5 - (double)cf.thr_u;             // eps + tie radius (>= eps)
        const double rel = fabs((double)Uf - ue) / band;
        if (rel > ratio[0]) ratio[0] = rel;
      }
  }
}

}  // extern "C"
