// TEST INFRASTRUCTURE ONLY -- never loaded by the meshflow_b200 package.
//
// Drives the __host__ __device__ per-element functions of meshflow_b200/csrc/mf_math.cuh on the CPU
// (g++ -ffp-contract=off, so no FMA contraction, like the explicit _rn intrinsics on the device) so
// that the CPU-only test-suite can compare the very arithmetic the kernels run against the oracle.
// It contains no kernels and no parallel plumbing; the GPU tests cover those.
#include <cstdint>
#include <cstring>
#include <vector>
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

}  // extern "C"
