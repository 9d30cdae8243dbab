// Mesh warp + crop on the device (hot-path subsystem 3).
//
// The reference loops over the R*C mesh cells of every frame and, for each cell, runs ~10 full-frame
// NumPy/OpenCV passes (mfs.py:1031-1061), then one cv2.remap (mfs.py:1063).  Here:
//
//   cell_setup_kernel : one thread per (frame, cell): the two exact 4-point homographies, the
//                       adjugate inverse used by the membership test, the integer bounds of the rest
//                       cell and the cell's support box in the output frame; the thread then files
//                       the cell into the candidate list of every 64x16 output tile its box touches.
//   warp_kernel       : one CTA per output tile.  Sorts the tile's candidates by descending cell id
//                       ("the last cell written wins", mfs.py:1060-1061), and per output pixel takes
//                       the first candidate whose warped rectangle mask is non-zero, evaluates the
//                       float32 remap coordinate in float64 exactly as the reference, gathers the
//                       four taps with OpenCV's 1/32-px fixed-point bilinear weights, and folds the
//                       four crop-edge searches (mfs.py:1075-1098) into warp reductions + atomics.
//                       Output rows are staged in shared memory and leave as 16-byte vectors.
//   crop_resize_kernel: cv2.resize of the cropped window back to W x H (mfs.py:1150-1155).
//
// HBM layout: frames [nf][H][W][3] uint8 (row pitch 3W, no padding -- BGR24 rows are multiples of 16
// bytes at every standard resolution); cells [nf][R*C] mf::Cell (168 B); tile lists
// [nf][tiles][kTileCap] uint16 + counts.
#include "mf_common.cuh"
#include "mf_math.cuh"

namespace mf {

static constexpr int kTileW = 64;
static constexpr int kTileH = 16;
static constexpr int kTileCap = 48;      // candidate cells per tile before the exhaustive fallback
static constexpr int kWarpThreads = 256; // 64 columns x 4 row phases

__global__ void __launch_bounds__(128) cell_setup_kernel(
    const double* __restrict__ u, const double* __restrict__ s, const float* __restrict__ vertex_xy,
    int nf, int W, int H, int R, int C, int tiles_x, int tiles_y, Cell* __restrict__ cells,
    int* __restrict__ tile_count, uint16_t* __restrict__ tile_list, int32_t* __restrict__ crop_out) {
  const int ncell = R * C;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nf * ncell) return;
  const int f = (int)(idx / ncell), id = (int)(idx - (int64_t)f * ncell);
  const int r = id / C, c = id - r * C;
  const int V = (R + 1) * (C + 1);
  if (id == 0 && crop_out) {  // identities of the max/max/min/min edge searches (mfs.py:992-995)
    crop_out[4 * f + 0] = 0; crop_out[4 * f + 1] = 0;
    crop_out[4 * f + 2] = W - 1; crop_out[4 * f + 3] = H - 1;
  }
  double rest[8], stab[8];
  const int vidx[4] = {r * (C + 1) + c, r * (C + 1) + c + 1, (r + 1) * (C + 1) + c, (r + 1) * (C + 1) + c + 1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int v = vidx[k];
    const size_t o = ((size_t)f * V + v) * 2;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const double rv = (double)vertex_xy[2 * v + d];
      rest[2 * k + d] = rv;
      // stabilized vertex = rest + (s - u), float64, then rounded to float32 by cv2.findHomography
      stab[2 * k + d] = (double)(float)MF_ADD(rv, MF_SUB(s[o + d], u[o + d]));   // mfs.py:964-967, 1025
    }
  }
  Cell cell;
  cell_setup(rest, stab, W, H, cell);
  cells[idx] = cell;
  if (cell.bx0 > cell.bx1) return;
  const int ntiles = tiles_x * tiles_y;
  const int tx0 = cell.bx0 / kTileW, tx1 = cell.bx1 / kTileW;
  const int ty0 = cell.by0 / kTileH, ty1 = cell.by1 / kTileH;
  for (int ty = ty0; ty <= ty1; ++ty) {
    for (int tx = tx0; tx <= tx1; ++tx) {
      const size_t t = (size_t)f * ntiles + (size_t)ty * tiles_x + tx;
      const int slot = atomicAdd(&tile_count[t], 1);
      if (slot < kTileCap) tile_list[t * kTileCap + slot] = (uint16_t)id;
    }
  }
}

__device__ __forceinline__ bool in_box(const Cell& c, int x, int y) {
  return x >= c.bx0 && x <= c.bx1 && y >= c.by0 && y <= c.by1;
}

template <bool kWriteMaps>
__global__ void __launch_bounds__(kWarpThreads) warp_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out,
    const Cell* __restrict__ cells, const int* __restrict__ tile_count,
    const uint16_t* __restrict__ tile_list, int32_t* __restrict__ crop_out, float* __restrict__ map_out,
    int W, int H, int ncell, int tiles_x, int tiles_y, int bb, int bg, int br) {
  __shared__ Cell s_cells[kTileCap];
  __shared__ int s_ids[kTileCap];
  __shared__ __align__(16) uint8_t s_out[kTileH][kTileW * 3];

  const int f = blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const size_t tile = (size_t)f * tiles_x * tiles_y + (size_t)blockIdx.y * tiles_x + blockIdx.x;
  const int tid = threadIdx.x;
  const Cell* fcells = cells + (size_t)f * ncell;
  const int nraw = tile_count[tile];
  const bool overflow = nraw > kTileCap;
  const int ncand = overflow ? 0 : nraw;

  if (!overflow) {
    // rank sort by descending id, then copy the candidates' parameters as 8-byte words
    int my_id = -1;
    if (tid < ncand) { my_id = tile_list[tile * kTileCap + tid]; s_ids[tid] = my_id; }
    __syncthreads();
    int rank = 0;
    if (tid < ncand)
      for (int j = 0; j < ncand; ++j) rank += (s_ids[j] > my_id) ? 1 : 0;
    __syncthreads();
    if (tid < ncand) s_ids[rank] = my_id;
    __syncthreads();
    constexpr int kWords = sizeof(Cell) / 8;
    for (int i = tid; i < ncand * kWords; i += kWarpThreads) {
      const int ci = i / kWords, wi = i - ci * kWords;
      reinterpret_cast<unsigned long long*>(&s_cells[ci])[wi] =
          reinterpret_cast<const unsigned long long*>(&fcells[s_ids[ci]])[wi];
    }
    __syncthreads();
  }

  const uint8_t* src = frames_in + (size_t)f * H * W * 3;
  const int lx = tid & (kTileW - 1), ly0 = tid >> 6;
  const int px = x0 + lx;
  int e_left = -1, e_top = -1, e_right = MF_INT_MAX, e_bottom = MF_INT_MAX;

#pragma unroll
  for (int j = 0; j < kTileH / 4; ++j) {
    const int ly = ly0 + 4 * j;
    const int py = y0 + ly;
    if (px < W && py < H) {
      const double x = (double)px, y = (double)py;
      float mx = (float)(W + 1), my = (float)(H + 1);          // mfs.py:983-984
      if (!overflow) {
        for (int k = 0; k < ncand; ++k) {
          const Cell& c = s_cells[k];
          if (in_box(c, px, py) && cell_inside(c, x, y)) { cell_map(c, x, y, mx, my); break; }
        }
      } else {
        for (int id = ncell - 1; id >= 0; --id) {
          const Cell& c = fcells[id];
          if (in_box(c, px, py) && cell_inside(c, x, y)) { cell_map(c, x, y, mx, my); break; }
        }
      }
      if (kWriteMaps) {
        float2* mo = reinterpret_cast<float2*>(map_out) + ((size_t)f * H + py) * W + px;
        *mo = make_float2(mx, my);
      }
      // crop-edge searches on the float32-valued map (mfs.py:1075-1098):  |m - e| < 1
      if (mx > -1.0f && mx < 1.0f) e_left = max(e_left, px);
      if (mx > (float)(W - 2) && mx < (float)W) e_right = min(e_right, px);
      if (my > -1.0f && my < 1.0f) e_top = max(e_top, py);
      if (my > (float)(H - 2) && my < (float)H) e_bottom = min(e_bottom, py);

      int ix, iy, ax, ay;
      remap_coords(mx, my, ix, iy, ax, ay);
      uint8_t* o = &s_out[ly][lx * 3];
      if (ix >= 0 && ix + 1 < W && iy >= 0 && iy + 1 < H) {
        const uint8_t* p0 = src + ((size_t)iy * W + ix) * 3;
        const uint8_t* p1 = p0 + (size_t)W * 3;
        const int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const int v = __ldg(p0 + ch) * w00 + __ldg(p0 + 3 + ch) * w01 + __ldg(p1 + ch) * w10 +
                        __ldg(p1 + 3 + ch) * w11;
          o[ch] = (uint8_t)((v + 512) >> 10);
        }
      } else {
        remap_pixel(src, W, H, ix, iy, ax, ay, bb, bg, br, o);
      }
    }
  }

  // crop edges: warp reduction, then at most four atomics per warp
  const unsigned full = 0xffffffffu;
  e_left = __reduce_max_sync(full, e_left);
  e_top = __reduce_max_sync(full, e_top);
  e_right = __reduce_min_sync(full, e_right);
  e_bottom = __reduce_min_sync(full, e_bottom);
  if ((tid & 31) == 0) {
    int32_t* cr = crop_out + 4 * f;
    if (e_left >= 0) atomicMax(cr + 0, e_left);
    if (e_top >= 0) atomicMax(cr + 1, e_top);
    if (e_right != MF_INT_MAX) atomicMin(cr + 2, e_right);
    if (e_bottom != MF_INT_MAX) atomicMin(cr + 3, e_bottom);
  }
  __syncthreads();

  uint8_t* dst = frames_out + (size_t)f * H * W * 3;
  const bool vec_ok = ((W * 3) % 16 == 0) && (x0 + kTileW <= W);
  if (vec_ok) {
    constexpr int kVecPerRow = kTileW * 3 / 16;  // 12
    for (int i = tid; i < kTileH * kVecPerRow; i += kWarpThreads) {
      const int row = i / kVecPerRow, v = i - row * kVecPerRow;
      if (y0 + row < H) {
        const uint4 val = *reinterpret_cast<const uint4*>(&s_out[row][v * 16]);
        *reinterpret_cast<uint4*>(dst + ((size_t)(y0 + row) * W + x0) * 3 + v * 16) = val;
      }
    }
  } else {
    const int wpx = min(kTileW, W - x0);
    for (int i = tid; i < kTileH * wpx * 3; i += kWarpThreads) {
      const int row = i / (wpx * 3), b = i - row * (wpx * 3);
      if (y0 + row < H) dst[((size_t)(y0 + row) * W + x0) * 3 + b] = s_out[row][b];
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

__global__ void __launch_bounds__(kWarpThreads) crop_resize_kernel(
    const uint8_t* __restrict__ frames_in, uint8_t* __restrict__ frames_out, int W, int H,
    const int32_t* __restrict__ enc4, const int4* __restrict__ xtab, const int4* __restrict__ ytab) {
  __shared__ __align__(16) uint8_t s_out[kTileH][kTileW * 3];
  int left, top, right_unused, bottom_unused;
  if (!decode_crop(enc4, W, H, left, top, right_unused, bottom_unused)) return;
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int tid = threadIdx.x;
  const int lx = tid & (kTileW - 1), ly0 = tid >> 6;
  const int px = x0 + lx;
  const uint8_t* src = frames_in + (size_t)f * H * W * 3;
  if (px < W) {
    const int4 xt = xtab[px];
    const int c0 = (left + xt.x) * 3, c1 = (left + xt.y) * 3;
#pragma unroll
    for (int j = 0; j < kTileH / 4; ++j) {
      const int ly = ly0 + 4 * j, py = y0 + ly;
      if (py < H) {
        const int4 yt = ytab[py];
        const uint8_t* r0 = src + (size_t)(top + yt.x) * W * 3;
        const uint8_t* r1 = src + (size_t)(top + yt.y) * W * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          s_out[ly][lx * 3 + ch] = (uint8_t)resize_blend(__ldg(r0 + c0 + ch), __ldg(r0 + c1 + ch),
                                                        __ldg(r1 + c0 + ch), __ldg(r1 + c1 + ch),
                                                        xt.z, xt.w, yt.z, yt.w);
        }
      }
    }
  }
  __syncthreads();
  uint8_t* dst = frames_out + (size_t)f * H * W * 3;
  const bool vec_ok = ((W * 3) % 16 == 0) && (x0 + kTileW <= W);
  if (vec_ok) {
    constexpr int kVecPerRow = kTileW * 3 / 16;
    for (int i = tid; i < kTileH * kVecPerRow; i += kWarpThreads) {
      const int row = i / kVecPerRow, v = i - row * kVecPerRow;
      if (y0 + row < H)
        *reinterpret_cast<uint4*>(dst + ((size_t)(y0 + row) * W + x0) * 3 + v * 16) =
            *reinterpret_cast<const uint4*>(&s_out[row][v * 16]);
    }
  } else {
    const int wpx = min(kTileW, W - x0);
    for (int i = tid; i < kTileH * wpx * 3; i += kWarpThreads) {
      const int row = i / (wpx * 3), b = i - row * (wpx * 3);
      if (y0 + row < H) dst[((size_t)(y0 + row) * W + x0) * 3 + b] = s_out[row][b];
    }
  }
}

struct WarpWorkspace {
  Cell* cells;
  int* tile_count;
  uint16_t* tile_list;
};

static bool carve_warp(Carver& cv, int nf, int W, int H, int R, int C, WarpWorkspace& w) {
  const size_t tiles = (size_t)((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  w.cells = cv.take<Cell>((size_t)nf * R * C);
  w.tile_count = cv.take<int>((size_t)nf * tiles);
  w.tile_list = cv.take<uint16_t>((size_t)nf * tiles * kTileCap);
  return cv.ok();
}

}  // namespace mf

extern "C" size_t mf_warp_workspace_bytes(int nf, int W, int H, int R, int C) {
  if (nf <= 0 || W <= 0 || H <= 0 || R <= 0 || C <= 0) return 0;
  mf::Carver cv(nullptr, 0);
  mf::WarpWorkspace w;
  mf::carve_warp(cv, nf, W, H, R, C, w);
  return mf::align_up(cv.used, 256);
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
  cudaError_t ce = cudaMemsetAsync(w.tile_count, 0, (size_t)nf * tiles_x * tiles_y * sizeof(int), st);
  if (ce != cudaSuccess) return mf::fail(MF_E_LAUNCH, "mf_warp_frames: memset: %s", cudaGetErrorString(ce));
  const int64_t ncells = (int64_t)nf * R * C;
  mf::cell_setup_kernel<<<(unsigned)((ncells + 127) / 128), 128, 0, st>>>(
      u, s, vertex_xy, nf, W, H, R, C, tiles_x, tiles_y, w.cells, w.tile_count, w.tile_list, crop_out);
  if (int e = mf::check_launch("cell_setup")) return e;
  const dim3 grid((unsigned)tiles_x, (unsigned)tiles_y, (unsigned)nf);
  if (map_out)
    mf::warp_kernel<true><<<grid, mf::kWarpThreads, 0, st>>>(frames_in, frames_out, w.cells, w.tile_count,
                                                            w.tile_list, crop_out, map_out, W, H, R * C,
                                                            tiles_x, tiles_y, border_b, border_g, border_r);
  else
    mf::warp_kernel<false><<<grid, mf::kWarpThreads, 0, st>>>(frames_in, frames_out, w.cells, w.tile_count,
                                                             w.tile_list, crop_out, nullptr, W, H, R * C,
                                                             tiles_x, tiles_y, border_b, border_g, border_r);
  return mf::check_launch("warp");
}

extern "C" size_t mf_crop_resize_workspace_bytes(int W, int H) {
  if (W <= 0 || H <= 0) return 0;
  return 256 + mf::align_up((size_t)W * sizeof(int4), 256) + mf::align_up((size_t)H * sizeof(int4), 256);
}

static int launch_crop_resize(const uint8_t* frames_in, int nf, int W, int H, const int32_t* enc4,
                              uint8_t* frames_out, void* workspace, cudaStream_t st) {
  int4* xtab = (int4*)((char*)workspace + 256);
  int4* ytab = (int4*)((char*)workspace + 256 + mf::align_up((size_t)W * sizeof(int4), 256));
  mf::resize_table_kernel<<<(W + H + 127) / 128, 128, 0, st>>>(W, H, enc4, xtab, ytab);
  if (int e = mf::check_launch("resize_table")) return e;
  const dim3 grid((unsigned)((W + mf::kTileW - 1) / mf::kTileW), (unsigned)((H + mf::kTileH - 1) / mf::kTileH),
                  (unsigned)nf);
  mf::crop_resize_kernel<<<grid, mf::kWarpThreads, 0, st>>>(frames_in, frames_out, W, H, enc4, xtab, ytab);
  return mf::check_launch("crop_resize");
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
