// HBM-/latency-bound helper kernels of the denoise step: embeddings, layout changes feeding the
// tensor-core kernels, the pose-encoder's small-channel convolutions, and the fused pipeline pieces
// (input assembly, CFG combine + per-frame DDIM update).  All global traffic is 128-bit where the
// layout allows it.
#include "kernels.h"

namespace d4d {

namespace {

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---------------------------------------------------------------------------------------------
// sinusoidal position embedding (upstream get_timestep_embedding; unet_multiview_condition.py:464,255)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void sinusoid_kernel(const T* __restrict__ pos, int n, int dim, int flip, float freq_shift,
                                bf16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= n * half) return;
  const int r = idx / half, i = idx % half;
  const float exponent = -logf(10000.0f) * static_cast<float>(i) / (static_cast<float>(half) - freq_shift);
  const float arg = static_cast<float>(pos[r]) * expf(exponent);
  const float s = sinf(arg), c = cosf(arg);
  bf16* o = out + static_cast<size_t>(r) * dim;
  if (flip) { o[i] = __float2bfloat16_rn(c); o[half + i] = __float2bfloat16_rn(s); }
  else { o[i] = __float2bfloat16_rn(s); o[half + i] = __float2bfloat16_rn(c); }
}

__global__ void silu_kernel(const bf16* __restrict__ x, long long n, bf16* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(silu_f(__bfloat162float(x[i])));
}

// ---------------------------------------------------------------------------------------------
// im2col of an NCHW input for the 3x3 pad-1 conv_in:  out[pixel, tap*cin_pad + c], zero padded to KP
// ---------------------------------------------------------------------------------------------
__global__ void im2col_nchw_kernel(const bf16* __restrict__ x, int n, int Cin, int H, int W, int cin_pad, int KP,
                                   bf16* __restrict__ out) {
  // one thread per (pixel, tap): writes cin_pad contiguous values
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int taps_total = KP / cin_pad;  // >= 9; taps >= 9 are zero padding
  const long long total = static_cast<long long>(n) * H * W * taps_total;
  if (idx >= total) return;
  const int tap = static_cast<int>(idx % taps_total);
  const long long pix = idx / taps_total;
  const int xw = static_cast<int>(pix % W);
  const int yh = static_cast<int>((pix / W) % H);
  const int img = static_cast<int>(pix / (static_cast<long long>(W) * H));
  bf16* o = out + pix * KP + tap * cin_pad;
  const int ky = tap / 3, kx = tap % 3;
  const int yy = yh + ky - 1, xx = xw + kx - 1;
  const bool in = tap < 9 && yy >= 0 && yy < H && xx >= 0 && xx < W;
  const bf16 zero = __float2bfloat16_rn(0.f);
  for (int c = 0; c < cin_pad; ++c) {
    bf16 v = zero;
    if (in && c < Cin) v = x[((static_cast<size_t>(img) * Cin + c) * H + yy) * W + xx];
    o[c] = v;
  }
}

// generic NHWC im2col (pad 1): out[opix, (ky*k+kx)*C + c]; used by the stride-2 downsample convs
// (Downsample2D, unet_multiview_blocks.py:460) and the pose encoder's 4x4 stride-2 conv
__global__ void im2col_nhwc_kernel(const bf16* __restrict__ x, int n, int H, int W, int C, int ksize, int stride,
                                   bf16* __restrict__ out) {
  const int Ho = (H + 2 - ksize) / stride + 1, Wo = (W + 2 - ksize) / stride + 1, oct = C / 8;
  const int kk = ksize * ksize;
  const long long total = static_cast<long long>(n) * Ho * Wo * kk * oct;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int o8 = static_cast<int>(idx % oct);
  long long r = idx / oct;
  const int tap = static_cast<int>(r % kk);
  r /= kk;
  const int xo = static_cast<int>(r % Wo);
  const int yo = static_cast<int>((r / Wo) % Ho);
  const int img = static_cast<int>(r / (static_cast<long long>(Wo) * Ho));
  const int yy = yo * stride + tap / ksize - 1, xx = xo * stride + tap % ksize - 1;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (yy >= 0 && yy < H && xx >= 0 && xx < W)
    v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(img) * H + yy) * W + xx) * C + o8 * 8));
  *reinterpret_cast<uint4*>(out + (r * kk + tap) * C + o8 * 8) = v;
}

__global__ void upsample2x_kernel(const bf16* __restrict__ x, int n, int H, int W, int C, bf16* __restrict__ out) {
  const int oct = C / 8;
  const long long total = static_cast<long long>(n) * (2 * H) * (2 * W) * oct;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int o8 = static_cast<int>(idx % oct);
  long long r = idx / oct;
  const int xo = static_cast<int>(r % (2 * W));
  const int yo = static_cast<int>((r / (2 * W)) % (2 * H));
  const int img = static_cast<int>(r / (static_cast<long long>(4) * W * H));
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(img) * H + yo / 2) * W + xo / 2) * C + o8 * 8));
  *reinterpret_cast<uint4*>(out + r * C + o8 * 8) = v;
}

// [n*H*W, ld] (first C columns) -> NCHW [n, C, H, W]
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, int ld, int n, int C, int hw, bf16* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * C * hw;
  if (idx >= total) return;
  const int p = static_cast<int>(idx % hw);
  const int c = static_cast<int>((idx / hw) % C);
  const int img = static_cast<int>(idx / (static_cast<long long>(hw) * C));
  out[idx] = x[(static_cast<size_t>(img) * hw + p) * ld + c];
}

// ---------------------------------------------------------------------------------------------
// direct small-channel convolution (pose encoder, pose_encoder.py:14-31): pad 1, k in {3,4}, stride {1,2}
// one thread per PIX horizontally adjacent output pixels, PIX x COUT accumulators in registers, weights [k*k][Cin][COUT] fp32
// in smem (every weight read from shared memory - a warp-wide broadcast - feeds PIX FMAs: with one pixel per thread the
// kernel was bound by those broadcasts, not by the FMAs)
// ---------------------------------------------------------------------------------------------
template <int COUT, int PIX>
__global__ void direct_conv_kernel(const bf16* __restrict__ x, int in_nchw, int n, int Cin, int H, int W,
                                   const bf16* __restrict__ w, const float* __restrict__ bias, int ksize, int stride,
                                   int silu, float out_scale, bf16* __restrict__ out) {
  extern __shared__ float sw[];  // [k*k*Cin][COUT]
  const int kk = ksize * ksize;
  for (int i = threadIdx.x; i < kk * Cin * COUT; i += blockDim.x) sw[i] = __bfloat162float(w[i]);
  __syncthreads();
  const int Ho = (H + 2 - ksize) / stride + 1, Wo = (W + 2 - ksize) / stride + 1;
  const int Wg = (Wo + PIX - 1) / PIX;  // pixel groups per row
  const long long total = static_cast<long long>(n) * Ho * Wg;
  const long long grp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (grp >= total) return;
  const int xo0 = static_cast<int>(grp % Wg) * PIX;
  const int yo = static_cast<int>((grp / Wg) % Ho);
  const int img = static_cast<int>(grp / (static_cast<long long>(Wg) * Ho));
  float acc[PIX][COUT];
#pragma unroll
  for (int p = 0; p < PIX; ++p)
#pragma unroll
    for (int i = 0; i < COUT; ++i) acc[p][i] = bias ? bias[i] : 0.f;
  for (int ky = 0; ky < ksize; ++ky) {
    const int yy = yo * stride + ky - 1;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < ksize; ++kx) {
      const float* wt = sw + static_cast<size_t>(ky * ksize + kx) * Cin * COUT;
      int xx[PIX];
      bool ok[PIX];
#pragma unroll
      for (int p = 0; p < PIX; ++p) {
        xx[p] = (xo0 + p) * stride + kx - 1;
        ok[p] = xx[p] >= 0 && xx[p] < W;  // (pixels past the row end read a clamped address and are not stored)
        if (!ok[p]) xx[p] = 0;
      }
      if (in_nchw) {
        for (int c = 0; c < Cin; ++c) {
          float v[PIX];
#pragma unroll
          for (int p = 0; p < PIX; ++p)
            v[p] = ok[p] ? __bfloat162float(x[((static_cast<size_t>(img) * Cin + c) * H + yy) * W + xx[p]]) : 0.f;
#pragma unroll
          for (int i = 0; i < COUT; ++i) {
            const float wv = wt[c * COUT + i];
#pragma unroll
            for (int p = 0; p < PIX; ++p) acc[p][i] = fmaf(v[p], wv, acc[p][i]);
          }
        }
      } else if (Cin % 8 == 0) {
        for (int c8 = 0; c8 < Cin; c8 += 8) {
          float f[PIX][8];
#pragma unroll
          for (int p = 0; p < PIX; ++p) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (ok[p]) u = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(img) * H + yy) * W + xx[p]) * Cin + c8));
            float2 q;
            q = unpack_bf16x2(u.x); f[p][0] = q.x; f[p][1] = q.y;
            q = unpack_bf16x2(u.y); f[p][2] = q.x; f[p][3] = q.y;
            q = unpack_bf16x2(u.z); f[p][4] = q.x; f[p][5] = q.y;
            q = unpack_bf16x2(u.w); f[p][6] = q.x; f[p][7] = q.y;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int i = 0; i < COUT; ++i) {
              const float wv = wt[(c8 + j) * COUT + i];
#pragma unroll
              for (int p = 0; p < PIX; ++p) acc[p][i] = fmaf(f[p][j], wv, acc[p][i]);
            }
          }
        }
      } else {
        for (int c = 0; c < Cin; ++c) {
          float v[PIX];
#pragma unroll
          for (int p = 0; p < PIX; ++p)
            v[p] = ok[p] ? __bfloat162float(x[((static_cast<size_t>(img) * H + yy) * W + xx[p]) * Cin + c]) : 0.f;
#pragma unroll
          for (int i = 0; i < COUT; ++i) {
            const float wv = wt[c * COUT + i];
#pragma unroll
            for (int p = 0; p < PIX; ++p) acc[p][i] = fmaf(v[p], wv, acc[p][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PIX; ++p) {
    if (xo0 + p >= Wo) break;
    bf16* o = out + ((static_cast<size_t>(img) * Ho + yo) * Wo + xo0 + p) * COUT;
#pragma unroll
    for (int i = 0; i < COUT; ++i) {
      float y = acc[p][i];
      if (silu) y = silu_f(y);
      o[i] = __float2bfloat16_rn(y * out_scale);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// a-1 input assembly (pipeline_diffuman4d.py:373-395)
// ---------------------------------------------------------------------------------------------
__global__ void assemble_kernel(const AssembleArgs a, int Cin) {
  const int f = blockIdx.y;
  const int hw = a.h * a.w;
  const bool is_cond = __bfloat162float(a.mask[static_cast<size_t>(f) * hw]) == 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long t = 0;
    if (!is_cond) {
      long long idx = a.timestep_indices[f];
      idx = idx < 0 ? 0 : (idx >= a.n_steps ? a.n_steps - 1 : idx);
      t = a.timesteps_table[idx];
    }
    a.timestep_out[f] = t;
    if (a.cfg) a.timestep_out[a.F + f] = t;
  }
  const bf16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f), mone = __float2bfloat16_rn(-1.f);
  const int halves = a.cfg ? 2 : 1;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const bf16 m = a.mask[static_cast<size_t>(f) * hw + p];
    for (int c = 0; c < 4; ++c) {
      const size_t li = (static_cast<size_t>(f) * 4 + c) * hw + p;
      bf16 lat = a.latents[li];
      if (is_cond) {
        lat = a.pixel[li];
        a.latents[li] = lat;  // reference aliasing quirk: latents <- image latents at cond frames (PIPE:375-379)
      }
      // positive half is the LAST half when cfg (torch.cat([negative, positive]))
      const int pos_img = a.cfg ? a.F + f : f;
      a.sample[(static_cast<size_t>(pos_img) * Cin + c) * hw + p] = lat;
      if (a.cfg) a.sample[(static_cast<size_t>(f) * Cin + c) * hw + p] = is_cond ? one : lat;
    }
    int ch = 4;
    for (int c = 0; c < 6; ++c, ++ch) {
      const bf16 v = a.plucker[(static_cast<size_t>(f) * 6 + c) * hw + p];
      const int pos_img = a.cfg ? a.F + f : f;
      a.sample[(static_cast<size_t>(pos_img) * Cin + ch) * hw + p] = v;
      if (a.cfg) a.sample[(static_cast<size_t>(f) * Cin + ch) * hw + p] = zero;
    }
    if (a.skel_latents) {
      for (int c = 0; c < 4; ++c, ++ch) {
        const bf16 v = a.skel_latents[(static_cast<size_t>(f) * 4 + c) * hw + p];
        const int pos_img = a.cfg ? a.F + f : f;
        a.sample[(static_cast<size_t>(pos_img) * Cin + ch) * hw + p] = v;
        if (a.cfg) a.sample[(static_cast<size_t>(f) * Cin + ch) * hw + p] = mone;
      }
    }
    for (int hf = 0; hf < halves; ++hf)
      a.sample[(static_cast<size_t>(hf * a.F + f) * Cin + ch) * hw + p] = m;
  }
}

// images 0..F-1 <- small image 0 ; images F..2F-1 <- small images 1..F   (per_img elements each, multiple of 8)
__global__ void broadcast_neg_images_kernel(const bf16* __restrict__ small, long long per_img8, int F, bf16* __restrict__ full) {
  const long long total = per_img8 * 2 * F;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long img = i / per_img8, off = i - img * per_img8;
    const long long src = img < F ? 0 : img - F + 1;
    reinterpret_cast<uint4*>(full)[i] = __ldg(reinterpret_cast<const uint4*>(small) + src * per_img8 + off);
  }
}
__global__ void fill_bf16_kernel(bf16* __restrict__ p, long long n, float v) {
  const bf16 b = __float2bfloat16_rn(v);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    p[i] = b;
}

__global__ void cfg_skeleton_kernel(const bf16* __restrict__ skel, long long per_frame, int F, bf16* __restrict__ out) {
  const long long total = per_frame * F;
  const bf16 mone = __float2bfloat16_rn(-1.f);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    out[i] = mone;
    out[total + i] = skel[i];
  }
}

// ---------------------------------------------------------------------------------------------
// a-13 + a-14: CFG combine + per-frame DDIM step (pipeline_diffuman4d.py:408-423, upstream DDIMScheduler.step)
// ---------------------------------------------------------------------------------------------
template <bool EMU>
__device__ __forceinline__ float rnd(float x) { return EMU ? bf16_round(x) : x; }

template <bool EMU>
__global__ void cfg_ddim_kernel(const DdimArgs a, long long* ts_out) {
  const int f = blockIdx.y;
  const bool is_cond = __bfloat162float(a.mask[static_cast<size_t>(f) * a.hw]) == 0.f;
  long long idx = a.timestep_indices[f];
  if (blockIdx.x == 0 && threadIdx.x == 0) ts_out[f] = is_cond ? 0 : idx + 1;
  const size_t base = static_cast<size_t>(f) * a.chw;
  if (is_cond) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.chw; i += gridDim.x * blockDim.x)
      a.out[base + i] = a.latents[base + i];
    return;
  }
  idx = idx < 0 ? 0 : (idx >= a.n_steps ? a.n_steps - 1 : idx);
  const long long t = a.timesteps_table[idx];
  const long long prev_t = t - a.T / a.n_steps;
  const float a_t = a.alphas_cumprod[t];
  const float a_prev = prev_t >= 0 ? a.alphas_cumprod[prev_t] : a.final_alpha_cumprod;
  const float b_t = 1.0f - a_t;
  const float sa = sqrtf(a_t), sb = sqrtf(b_t);
  const float sa_prev = sqrtf(a_prev), sdir = sqrtf(1.0f - a_prev);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.chw; i += gridDim.x * blockDim.x) {
    float eps_in;
    if (a.cfg) {
      const float u = __bfloat162float(a.noise[base + i]);
      const float c = __bfloat162float(a.noise[static_cast<size_t>(a.F) * a.chw + base + i]);
      // u + g * (c - u)
      eps_in = rnd<EMU>(u + rnd<EMU>(a.guidance * rnd<EMU>(c - u)));
    } else {
      eps_in = __bfloat162float(a.noise[base + i]);
    }
    const float x = __bfloat162float(a.latents[base + i]);
    float x0, eps;
    if (a.prediction_type == 0) {        // epsilon
      x0 = rnd<EMU>(rnd<EMU>(x - rnd<EMU>(sb * eps_in)) / sa);
      eps = eps_in;
    } else if (a.prediction_type == 1) { // v_prediction
      x0 = rnd<EMU>(rnd<EMU>(sa * x) - rnd<EMU>(sb * eps_in));
      eps = rnd<EMU>(rnd<EMU>(sa * eps_in) + rnd<EMU>(sb * x));
    } else {                             // sample
      x0 = eps_in;
      eps = rnd<EMU>(rnd<EMU>(x - rnd<EMU>(sa * x0)) / sb);
    }
    if (a.clip_sample) x0 = fminf(fmaxf(x0, -a.clip_range), a.clip_range);
    const float dir = rnd<EMU>(sdir * eps);
    const float prev = rnd<EMU>(rnd<EMU>(sa_prev * x0) + dir);
    a.out[base + i] = __float2bfloat16_rn(prev);
  }
}

// ---------------------------------------------------------------------------------------------
// frame-sharded window: K/V arrival flags in peer memory
// ---------------------------------------------------------------------------------------------
__global__ void kv_signal_kernel(const KvFlagArgs a) {
  const int r = threadIdx.x;
  if (r >= a.world) return;
  __threadfence_system();  // order the K/V stores of the preceding kernels (this stream) before the flag
  unsigned int* f = a.flags[r] + a.slot * 8 + a.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(a.epoch) : "memory");
}
__global__ void kv_wait_kernel(const KvFlagArgs a) {
  const int r = threadIdx.x;
  if (r >= a.world) return;
  const unsigned int* f = a.flags[a.rank] + a.slot * 8 + r;
  unsigned int v = 0;
  unsigned long long spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (++spins > (1ull << 31)) {  // ~ seconds: a lost peer must trap, not hang the box
      printf("d4d: K/V flag timeout rank=%d waiting for rank=%d epoch=%u have=%u\n", a.rank, r, a.epoch, v);
      __trap();
    }
  } while (static_cast<int>(v - a.epoch) < 0);
  __threadfence_system();
}

inline int blocks_for(long long total, int threads) { return static_cast<int>((total + threads - 1) / threads); }

}  // namespace

int sinusoid_run(const float* pos, int n, int dim, int flip, float freq_shift, bf16* out, cudaStream_t stream) {
  D4D_REQUIRE(dim % 2 == 0 && n > 0, "sinusoid dims");
  const int total = n * (dim / 2);
  sinusoid_kernel<float><<<blocks_for(total, 256), 256, 0, stream>>>(pos, n, dim, flip, freq_shift, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int sinusoid_i64_run(const long long* pos, int n, int dim, int flip, float freq_shift, bf16* out, cudaStream_t stream) {
  D4D_REQUIRE(dim % 2 == 0 && n > 0, "sinusoid dims");
  const int total = n * (dim / 2);
  sinusoid_kernel<long long><<<blocks_for(total, 256), 256, 0, stream>>>(pos, n, dim, flip, freq_shift, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int silu_run(const bf16* x, long long n, bf16* out, cudaStream_t stream) {
  if (n <= 0) return 0;
  silu_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(x, n, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int im2col_nchw_run(const bf16* x, int n, int Cin, int H, int W, int cin_pad, int KP, bf16* out, cudaStream_t stream) {
  D4D_REQUIRE(cin_pad >= Cin && KP % cin_pad == 0 && KP >= 9 * cin_pad, "im2col padding");
  const long long total = static_cast<long long>(n) * H * W * (KP / cin_pad);
  im2col_nchw_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(x, n, Cin, H, W, cin_pad, KP, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int im2col_nhwc_run(const bf16* x, int n, int H, int W, int C, int ksize, int stride, bf16* out, cudaStream_t stream) {
  D4D_REQUIRE(C % 8 == 0 && (ksize == 3 || ksize == 4) && (stride == 1 || stride == 2), "im2col_nhwc dims");
  const int Ho = (H + 2 - ksize) / stride + 1, Wo = (W + 2 - ksize) / stride + 1;
  const long long total = static_cast<long long>(n) * Ho * Wo * ksize * ksize * (C / 8);
  im2col_nhwc_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(x, n, H, W, C, ksize, stride, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int upsample2x_run(const bf16* x, int n, int H, int W, int C, bf16* out, cudaStream_t stream) {
  D4D_REQUIRE(C % 8 == 0, "upsample channels");
  const long long total = static_cast<long long>(n) * 4 * H * W * (C / 8);
  upsample2x_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(x, n, H, W, C, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int nhwc_to_nchw_run(const bf16* x, int ld, int n, int C, int hw, bf16* out, cudaStream_t stream) {
  const long long total = static_cast<long long>(n) * C * hw;
  nhwc_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(x, ld, n, C, hw, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int direct_conv_run(const bf16* x, int in_nchw, int n, int Cin, int H, int W, const bf16* w, const float* bias, int Cout,
                    int ksize, int stride, int silu, float out_scale, bf16* out_nhwc, cudaStream_t stream) {
  D4D_REQUIRE(ksize == 3 || ksize == 4, "direct conv kernel size");
  D4D_REQUIRE(stride == 1 || stride == 2, "direct conv stride");
  const int Ho = (H + 2 - ksize) / stride + 1, Wo = (W + 2 - ksize) / stride + 1;
  const long long total = static_cast<long long>(n) * Ho * Wo;
  const size_t smem = sizeof(float) * ksize * ksize * Cin * Cout;
  D4D_REQUIRE(smem <= 48 * 1024, "direct conv weights exceed 48 KB of shared memory");
  const int threads = 128;
  // pixels per thread: 4 for few output channels, 2 for 32 (64 accumulators), 1 for 64
#define D4D_DC(CO, PX)                                                                                                 \
  case CO: {                                                                                                           \
    const long long groups = static_cast<long long>(n) * Ho * ((Wo + PX - 1) / PX);                                     \
    direct_conv_kernel<CO, PX><<<blocks_for(groups, threads), threads, smem, stream>>>(x, in_nchw, n, Cin, H, W, w, bias, \
                                                                                       ksize, stride, silu, out_scale,  \
                                                                                       out_nhwc);                       \
  } break;
  switch (Cout) {
    D4D_DC(3, 4)
    D4D_DC(16, 4)
    D4D_DC(32, 2)
    D4D_DC(64, 1)
    default:
      set_error("direct conv: unsupported Cout " + std::to_string(Cout));
      return 1;
  }
#undef D4D_DC
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int kv_signal_run(const KvFlagArgs& a, cudaStream_t stream) {
  kv_signal_kernel<<<1, 32, 0, stream>>>(a);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}
int kv_wait_run(const KvFlagArgs& a, cudaStream_t stream) {
  kv_wait_kernel<<<1, 32, 0, stream>>>(a);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int assemble_input_run(const AssembleArgs& a, cudaStream_t stream) {
  const int Cin = 4 + 6 + (a.skel_latents ? 4 : 0) + 1;
  const int hw = a.h * a.w;
  dim3 grid(min(64, blocks_for(hw, 256)), a.F);
  assemble_kernel<<<grid, 256, 0, stream>>>(a, Cin);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int broadcast_neg_images_run(const bf16* small, long long per_img, int F, bf16* full, cudaStream_t stream) {
  D4D_REQUIRE(per_img % 8 == 0 && F > 0, "broadcast_neg_images arguments");
  broadcast_neg_images_kernel<<<148 * 8, 256, 0, stream>>>(small, per_img / 8, F, full);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}
int fill_bf16_run(bf16* p, long long n, float v, cudaStream_t stream) {
  fill_bf16_kernel<<<148 * 4, 256, 0, stream>>>(p, n, v);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int cfg_skeleton_run(const bf16* skel, long long per_frame_elems, int F, bf16* out, cudaStream_t stream) {
  const long long total = per_frame_elems * F;
  long long nb = (total + 255) / 256;
  const int blocks = static_cast<int>(nb < 148 * 16 ? nb : 148 * 16);
  cfg_skeleton_kernel<<<blocks, 256, 0, stream>>>(skel, per_frame_elems, F, out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

int cfg_ddim_step_run(const DdimArgs& a, long long* ts_out, cudaStream_t stream) {
  D4D_REQUIRE(a.n_steps > 0 && a.T > 0 && a.F > 0, "ddim args");
  dim3 grid(min(64, blocks_for(a.chw, 256)), a.F);
  if (a.emulate_bf16) cfg_ddim_kernel<true><<<grid, 256, 0, stream>>>(a, ts_out);
  else cfg_ddim_kernel<false><<<grid, 256, 0, stream>>>(a, ts_out);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace d4d
