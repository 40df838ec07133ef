// mg_wrappers.cu — SURVEY 8(f-3): the reference's observation wrappers (minigrid/wrappers.py) for a whole batch, on the
// device, so that no per-step Python post-processing is left on a training loop:
//   ViewSizeWrapper          :629-673   gen_obs with another (odd) agent_view_size          k_view
//   OneHotPartialObsWrapper  :217-284   (type, colour, state) -> 11 + 6 + 3 one-hot bytes   k_onehot
//   FlatObsWrapper           :557-626   image bytes ++ one-hot mission characters           k_flat
//   SymbolicObsWrapper       :729-782   (x, y, type | -1) int64 per cell                    k_symbolic
//   RGBImgPartialObsWrapper  :334-380   the agent's view rendered with 8 x 8 tiles          k_rgb_partial
//   RGBImgObsWrapper         :287-331   the whole grid rendered, agent's view highlighted   k_rgb_full
// These are consumers of K1's outputs or of the state arena, off the step path proper: plain kernels, one pass over
// their output, bound by it. The tile atlas of the two RGB wrappers is data rendered once by the reference's own
// Grid.render_tile (scripts/bake_tile_atlas.py) and handed in by the host layer.
#include "mg_common.cuh"

namespace mg {

__device__ __forceinline__ uint32_t code_at(const Params &p, int env, int x, int y) {
  if (x < 0 || y < 0 || x >= p.g.W || y >= p.g.H) return CODE_WALL;  // Grid.slice: outside the grid is grey wall (grid.py:136-139)
  return reinterpret_cast<const uint8_t *>(p.grid)[cell_byte_C(p.g, env, x, y)];
}

constexpr int MAX_VIEW = 15;

// The agent's view for any odd V <= 15: codes[vx * V + vy] (0 = not visible) exactly as MiniGridEnv.gen_obs_grid +
// Grid.process_vis produce it (minigrid_env.py:597-630, grid.py:291-328), carried object at (V / 2, V - 1).
// World cell of view cell (vx, vy): agent + d * (V - 1 - vy) + r * (vx - V / 2), d = DIR_TO_VEC[dir], r = (-d.y, d.x)
// (the closed form of get_view_exts + slice + (dir + 1) x rotate_left).
__device__ void view_codes(const Params &p, int env, int V, uint8_t *codes) {
  const uint4 rec = p.agent[env];
  const int ax = rec.x & 0xFF, ay = (rec.x >> 8) & 0xFF, dir = rec.y & 3;
  const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);
  const int rx = -dy, ry = dx;
  uint32_t opaque[MAX_VIEW];  // bit vx of row vy
  for (int vy = 0; vy < V; ++vy) {
    uint32_t op = 0;
    for (int vx = 0; vx < V; ++vx) {
      const int f = V - 1 - vy, l = vx - V / 2;
      const uint32_t c = code_at(p, env, ax + dx * f + rx * l, ay + dy * f + ry * l);
      codes[vx * V + vy] = (uint8_t)c;
      op |= ((c >> 7) & 1u) << vx;
    }
    opaque[vy] = op;
  }
  if (!p.see_through) {
    uint32_t mask[MAX_VIEW];
    for (int j = 0; j < V; ++j) mask[j] = 0;
    mask[V - 1] = 1u << (V / 2);
    for (int j = V - 1; j >= 0; --j) {
      for (int i = 0; i < V - 1; ++i) {
        if (!((mask[j] >> i) & 1u) || ((opaque[j] >> i) & 1u)) continue;
        mask[j] |= 1u << (i + 1);
        if (j > 0) mask[j - 1] |= (1u << (i + 1)) | (1u << i);
      }
      for (int i = V - 1; i >= 1; --i) {
        if (!((mask[j] >> i) & 1u) || ((opaque[j] >> i) & 1u)) continue;
        mask[j] |= 1u << (i - 1);
        if (j > 0) mask[j - 1] |= (1u << (i - 1)) | (1u << i);
      }
    }
    for (int vy = 0; vy < V; ++vy)
      for (int vx = 0; vx < V; ++vx)
        if (!((mask[vy] >> vx) & 1u)) codes[vx * V + vy] = 0;
  }
  codes[(V / 2) * V + V - 1] = (uint8_t)(rec.z ? rec.z : CODE_EMPTY);  // minigrid_env.py:623-630
}

__global__ void __launch_bounds__(128) k_view(Params p, int V, uint8_t *__restrict__ out) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= p.n_envs) return;
  uint8_t codes[MAX_VIEW * MAX_VIEW];
  view_codes(p, env, V, codes);
  uint8_t *o = out + (size_t)env * V * V * 3;
  for (int q = 0; q < V * V; ++q) {
    const uint32_t t = __ldg(p.cell_lut + codes[q]);
    o[3 * q] = (uint8_t)t; o[3 * q + 1] = (uint8_t)(t >> 8); o[3 * q + 2] = (uint8_t)(t >> 16);
  }
}

// one thread per cell: 3 bytes in, 20 bytes (5 aligned words) out
__global__ void k_onehot(const uint8_t *__restrict__ img, uint8_t *__restrict__ out, long long n_cells) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint32_t ch[3] = {(uint32_t)img[3 * c], 11u + img[3 * c + 1], 17u + img[3 * c + 2]};
  uint32_t w[5] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if ((ch[k] >> 2) == (uint32_t)j) w[j] |= 1u << (8 * (ch[k] & 3u));
  uint32_t *o = reinterpret_cast<uint32_t *>(out) + 5 * c;
#pragma unroll
  for (int j = 0; j < 5; ++j) o[j] = w[j];
}

// out[e] = img[e] (img_bytes) ++ mission (mission_bytes), one thread per output byte
__global__ void k_flat(const uint8_t *__restrict__ img, const uint8_t *__restrict__ mission, uint8_t *__restrict__ out, int img_bytes,
                       int mission_bytes, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int row = img_bytes + mission_bytes;
  const long long e = i / row;
  const int j = (int)(i - e * row);
  out[i] = j < img_bytes ? img[e * img_bytes + j] : mission[j - img_bytes];
}

__global__ void k_symbolic(Params p, long long *__restrict__ out) {
  const int cells = p.g.W * p.g.H;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.n_envs * cells) return;
  const int env = (int)(gid / cells), c = (int)(gid % cells);
  const int x = c / p.g.H, y = c % p.g.H;
  const uint32_t code = reinterpret_cast<const uint8_t *>(p.grid)[cell_byte_C(p.g, env, x, y)];
  const uint32_t t4 = code & 15u;
  long long type = t4 <= T_EMPTY ? -1 : (t4 == T4_BOX_WITH_KEY ? (long long)T_BOX : (t4 >= T4_DOOR_CLOSED ? (long long)T_DOOR : (long long)t4));
  const uint4 rec = p.agent[env];
  if ((int)(rec.x & 0xFF) == x && (int)((rec.x >> 8) & 0xFF) == y) type = T_AGENT;
  long long *o = out + gid * 3;
  o[0] = x; o[1] = y; o[2] = type;
}

// ---- RGB rendering from the tile atlas ----
// atlas_index[code (7 bits)][agent: 0 none, 1 + dir][highlight] -> tile number, tiles[t][8][8][3] (rendered by the reference)
constexpr int TILE_PX = 8, TILE_BYTES = TILE_PX * TILE_PX * 3;
__device__ __forceinline__ int atlas_tile(const uint16_t *index, uint32_t code, int agent, int highlight) {
  return index[((code & 0x7Fu) * 5 + agent) * 2 + highlight];
}
// copies row `py` of a tile (24 bytes = 6 words) to dst
__device__ __forceinline__ void put_tile_row(uint8_t *dst, const uint8_t *tiles, int tile, int py) {
  const uint32_t *s = reinterpret_cast<const uint32_t *>(tiles + (size_t)tile * TILE_BYTES + py * TILE_PX * 3);
  uint32_t *d = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = __ldg(s + k);
}

// RGBImgPartialObsWrapper: img[n][7][7][3] (K1's observation) -> out[n][56][56][3]. get_pov_render = the view grid (cells
// that are not visible were set to None by process_vis) rendered with highlight = vis_mask and the agent at (3, 6) facing
// up (dir 3) over what it carries (minigrid_env.py:652-676, grid.py:200-242). One thread per (env, cell, pixel row).
__global__ void k_rgb_partial(const uint8_t *__restrict__ img, const uint8_t *__restrict__ tiles, const uint16_t *__restrict__ index,
                              uint8_t *__restrict__ out, long long n_envs) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_envs * VIEW * VIEW * TILE_PX) return;
  const int py = (int)(gid % TILE_PX);
  const int cell = (int)((gid / TILE_PX) % (VIEW * VIEW));
  const long long env = gid / (TILE_PX * VIEW * VIEW);
  const int vy = cell / VIEW, vx = cell % VIEW;  // consecutive threads: consecutive pixel rows, then cells along x
  const uint8_t *c3 = img + (env * VIEW * VIEW + vx * VIEW + vy) * 3;
  const uint32_t type = c3[0], color = c3[1], state = c3[2];
  const bool seen = type != T_UNSEEN;
  const uint32_t code = seen ? encode_cell(type, color, state) : CODE_EMPTY;
  const int agent = (vx == VIEW / 2 && vy == VIEW - 1) ? 1 + 3 : 0;
  uint8_t *dst = out + ((env * VIEW * TILE_PX + vy * TILE_PX + py) * (VIEW * TILE_PX) + vx * TILE_PX) * 3;
  put_tile_row(dst, tiles, atlas_tile(index, code, agent, seen ? 1 : 0), py);
}

// RGBImgObsWrapper: the whole grid, out[n][H * 8][W * 8][3]; highlight = the cells of the agent's view that process_vis
// marks visible (get_full_render, minigrid_env.py:678-742). `vis` is this step's observation image (a cell of the view is
// visible iff its type is not 0): img[n][7][7][3]. One thread per (env, cell, pixel row).
__global__ void k_rgb_full(Params p, const uint8_t *__restrict__ img, const uint8_t *__restrict__ tiles, const uint16_t *__restrict__ index,
                           uint8_t *__restrict__ out) {
  const int W = p.g.W, H = p.g.H;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.n_envs * W * H * TILE_PX) return;
  const int py = (int)(gid % TILE_PX);
  const int cell = (int)((gid / TILE_PX) % (W * H));
  const int env = (int)(gid / ((long long)TILE_PX * W * H));
  const int y = cell / W, x = cell % W;
  const uint4 rec = p.agent[env];
  const int ax = rec.x & 0xFF, ay = (rec.x >> 8) & 0xFF, dir = rec.y & 3;
  // view coordinates of (x, y): inverse of world = agent + d * (6 - vy) + r * (vx - 3)
  const int dx = (dir == 0) - (dir == 2), dy = (dir == 1) - (dir == 3);
  const int ox = x - ax, oy = y - ay;
  const int f = ox * dx + oy * dy, l = ox * (-dy) + oy * dx;
  const int vx = l + VIEW / 2, vy = VIEW - 1 - f;
  int highlight = 0;
  if (vx >= 0 && vx < VIEW && vy >= 0 && vy < VIEW) highlight = img[((size_t)env * VIEW * VIEW + vx * VIEW + vy) * 3] != T_UNSEEN;
  const uint32_t code = reinterpret_cast<const uint8_t *>(p.grid)[cell_byte_C(p.g, env, x, y)];
  const int agent = (x == ax && y == ay) ? 1 + dir : 0;
  uint8_t *dst = out + (((size_t)env * H * TILE_PX + (size_t)y * TILE_PX + py) * (W * TILE_PX) + (size_t)x * TILE_PX) * 3;
  put_tile_row(dst, tiles, atlas_tile(index, code, agent, highlight), py);
}

cudaError_t launch_view(const Params &p, int V, uint8_t *out, cudaStream_t s) {
  k_view<<<(p.n_envs + 127) / 128, 128, 0, s>>>(p, V, out);
  return cudaGetLastError();
}
cudaError_t launch_onehot(const uint8_t *img, uint8_t *out, long long n_cells, cudaStream_t s) {
  k_onehot<<<(unsigned)((n_cells + 255) / 256), 256, 0, s>>>(img, out, n_cells);
  return cudaGetLastError();
}
cudaError_t launch_flat(const uint8_t *img, const uint8_t *mission, uint8_t *out, int img_bytes, int mission_bytes, long long n_envs,
                        cudaStream_t s) {
  const long long total = n_envs * (img_bytes + mission_bytes);
  k_flat<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(img, mission, out, img_bytes, mission_bytes, total);
  return cudaGetLastError();
}
cudaError_t launch_symbolic(const Params &p, long long *out, cudaStream_t s) {
  const long long total = (long long)p.n_envs * p.g.W * p.g.H;
  k_symbolic<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p, out);
  return cudaGetLastError();
}
cudaError_t launch_rgb_partial(const uint8_t *img, const uint8_t *tiles, const uint16_t *index, uint8_t *out, long long n_envs, cudaStream_t s) {
  const long long total = n_envs * VIEW * VIEW * TILE_PX;
  k_rgb_partial<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(img, tiles, index, out, n_envs);
  return cudaGetLastError();
}
cudaError_t launch_rgb_full(const Params &p, const uint8_t *img, const uint8_t *tiles, const uint16_t *index, uint8_t *out, cudaStream_t s) {
  const long long total = (long long)p.n_envs * p.g.W * p.g.H * TILE_PX;
  k_rgb_full<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p, img, tiles, index, out);
  return cudaGetLastError();
}

}  // namespace mg
