// Foot-contact classifier of contact-human-dynamics on sm_100a (product code).
//
// Replaces, for inference, src/contact_learning/test.py:51-152 (val_full_video) +
// src/contact_learning/models/openpose_only.py:29-78 + the window construction of
// src/contact_learning/data/real_video_dataset.py:206-276:
//   chd_k_contact_mlp  : gathers the 9-frame x 13-joint x (x,y,conf) windows straight from the per-frame keypoints
//                        (root-relative as the dataset does it, fp64 subtraction then fp32), and runs the whole MLP
//                        351-1024-512-128-32-20 (Linear + eval BatchNorm + ReLU) with activations in shared memory
//   chd_k_contact_vote : sigmoid > 0.5, 5-vote aggregation, edge thresholds, 2-frame padding, int64 labels
// fp32 FFMA with fp32 accumulation (no tf32/bf16): the labels must match the reference's fp32 forward bit for bit
// away from the logit-0 boundary (SURVEY 8(a)-D note).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/chd.h"

#define CT_TM 32        // windows per CTA
#define CT_THREADS 256
#define CT_WIN 9
#define CT_PRED 5
#define CT_J 13
#define CT_IN 351

static __device__ __constant__ int c_lower_joints[CT_J] = {8, 9, 10, 11, 12, 13, 14, 19, 20, 21, 22, 23, 24};  // openpose_dataset.py:38

struct ContactDev {
  const float* W[5];   // [in][out] (transposed on the host)
  const float* b[5];
  const float* bn_scale[4];  // gamma / sqrt(var + eps)
  const float* bn_mean[4];
  const float* bn_beta[4];
};

// out[m][n] = relu( bn( in[m][:] . W[:][n] + b[n] ) ) for the CTA's CT_TM windows.
// warp w: rows (w & 3)*8 .. +7, column half (w >> 2); per pass 4 columns per lane strided by 32 (coalesced weights).
template <bool BN_RELU>
__device__ __forceinline__ void contact_layer(const float* __restrict__ in, int Kp, int K, const float* __restrict__ W,
                                              const float* __restrict__ bias, const float* __restrict__ scale,
                                              const float* __restrict__ mean, const float* __restrict__ beta, int N,
                                              float* __restrict__ out, int Np) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mg = warp & 3, nh = warp >> 2;
  const float* a0 = in + (size_t)(mg * 8) * Kp;
  for (int n0 = nh * 128; n0 < N; n0 += 256) {
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    int nc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) nc[j] = n0 + lane + 32 * j;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float wv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = nc[j] < N ? __ldg(W + (size_t)k * N + nc[j]) : 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float a = a0[r * Kp + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = fmaf(a, wv[j], acc[r][j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (nc[j] >= N) continue;
      const float bj = bias[nc[j]];
      float sc = 1.f, mu = 0.f, be = 0.f;
      if (BN_RELU) sc = scale[nc[j]], mu = mean[nc[j]], be = beta[nc[j]];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float y = acc[r][j] + bj;
        if (BN_RELU) {
          y = (y - mu) * sc + be;
          y = fmaxf(y, 0.f);
        }
        out[(size_t)(mg * 8 + r) * Np + nc[j]] = y;
      }
    }
  }
}

// frames: [V][Fmax][25][3] fp64 (scaled, gap-interpolated, normalised keypoints); logits: [V*Wn][20]
__global__ void __launch_bounds__(CT_THREADS) chd_k_contact_mlp(ContactDev net, const double* __restrict__ frames, int V, int Fmax,
                                                                float* __restrict__ logits) {
  extern __shared__ float smf[];
  float* bufA = smf;                 // 32 x 1024
  float* bufB = smf + CT_TM * 1024;  // 32 x 512
  const int Wn = Fmax - (CT_WIN - 1);
  const int total = V * Wn;
  const int g0 = blockIdx.x * CT_TM;
  // gather windows into bufB [32][352] (real_video_dataset.py:240-252)
  for (int idx = threadIdx.x; idx < CT_TM * 352; idx += blockDim.x) {
    const int m = idx / 352, k = idx % 352;
    float val = 0.f;
    const int g = g0 + m;
    if (g < total && k < CT_IN) {
      const int v = g / Wn, w = g % Wn;
      const int f = k / (CT_J * 3), rem = k % (CT_J * 3), j = rem / 3, c = rem % 3;
      const int joint = c_lower_joints[j];
      const double* fr = frames + (((size_t)v * Fmax + w + f) * 25 + joint) * 3;
      if (c == 2) {
        val = (float)fr[2];
      } else {
        const double root = frames[(((size_t)v * Fmax + w + CT_WIN / 2) * 25 + 8) * 3 + c];
        val = (f == CT_WIN / 2 && joint == 8) ? (float)root : (float)(fr[c] - root);
      }
    }
    bufB[idx] = val;
  }
  __syncthreads();
  contact_layer<true>(bufB, 352, CT_IN, net.W[0], net.b[0], net.bn_scale[0], net.bn_mean[0], net.bn_beta[0], 1024, bufA, 1024);
  __syncthreads();
  contact_layer<true>(bufA, 1024, 1024, net.W[1], net.b[1], net.bn_scale[1], net.bn_mean[1], net.bn_beta[1], 512, bufB, 512);
  __syncthreads();
  contact_layer<true>(bufB, 512, 512, net.W[2], net.b[2], net.bn_scale[2], net.bn_mean[2], net.bn_beta[2], 128, bufA, 128);
  __syncthreads();
  contact_layer<true>(bufA, 128, 128, net.W[3], net.b[3], net.bn_scale[3], net.bn_mean[3], net.bn_beta[3], 32, bufB, 32);
  __syncthreads();
  contact_layer<false>(bufB, 32, 32, net.W[4], net.b[4], nullptr, nullptr, nullptr, 20, bufA, 20);
  __syncthreads();
  for (int idx = threadIdx.x; idx < CT_TM * 20; idx += blockDim.x) {
    const int g = g0 + idx / 20;
    if (g < total) logits[(size_t)g * 20 + idx % 20] = bufA[idx];
  }
}

// labels [V][Fmax][4] int64; rows >= seq_len are zeroed (the reference trims them, test.py:149-152)
__global__ void chd_k_contact_vote(const float* __restrict__ logits, int V, int Fmax, const int* __restrict__ seq_lens,
                                   long long* __restrict__ labels, float* __restrict__ min_abs) {
  const int Wn = Fmax - (CT_WIN - 1), nv = Wn + 2 * (CT_PRED / 2);   // frames that receive votes
  const int off = (CT_WIN - CT_PRED) / 2;                            // copies padded on each side
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float mabs = 3.4e38f;
  if (idx < V * Fmax * 4) {
    const int c = idx & 3, f = (idx >> 2) % Fmax, v = (idx >> 2) / Fmax;
    int fv = f - off;                        // index into the voted array, clamped = repeat first/last row
    fv = fv < 0 ? 0 : (fv >= nv ? nv - 1 : fv);
    int votes = 0;
    for (int p = 0; p < CT_PRED; ++p) {      // window w = fv - p contributes its prediction for offset p
      const int w = fv - p;
      if (w < 0 || w >= Wn) continue;
      const float x = logits[((size_t)v * Wn + w) * 20 + p * 4 + c];
      const float prob = 1.0f / (1.0f + expf(-x));   // openpose_only.py:75-78: sigmoid(x) > 0.5
      votes += prob > 0.5f ? 1 : 0;
      mabs = fminf(mabs, fabsf(x));
    }
    int thresh = (CT_PRED + 1) / 2;          // test.py:101-104
    const int e0 = fv, e1 = nv - 1 - fv;
    if (e0 < CT_PRED - 1) thresh = e0 / 2 + 1;
    if (e1 < CT_PRED - 1) thresh = e1 / 2 + 1;
    labels[idx] = f < seq_lens[v] ? (votes >= thresh ? 1 : 0) : 0;
  }
  // block min of |logit|
  for (int o = 16; o > 0; o >>= 1) mabs = fminf(mabs, __shfl_xor_sync(0xffffffffu, mabs, o));
  if ((threadIdx.x & 31) == 0 && mabs < 3.0e38f) atomicMin(reinterpret_cast<int*>(min_abs), __float_as_int(mabs));  // positive floats order as ints
}

struct chd_contact_net {
  std::vector<void*> allocs;
  ContactDev dev;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
};

#define CT_CUDA(x)                                                                           \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) {                                                                 \
      fprintf(stderr, "libchd: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -100 - (int)e_;                                                                 \
    }                                                                                        \
  } while (0)

extern "C" {

int chd_contact_create(const float* weights, const float* biases, const float* bn, float bn_eps, int32_t device, chd_contact_net** out) {
  if (!weights || !biases || !bn || !out) return -1;
  if (device >= 0) CT_CUDA(cudaSetDevice(device));
  const int dims[6] = {CT_IN, 1024, 512, 128, 32, 20};
  chd_contact_net* net = new chd_contact_net();
  CT_CUDA(cudaStreamCreate(&net->stream));
  auto up = [&](const std::vector<float>& h, const float** d) -> int {
    void* p = nullptr;
    CT_CUDA(cudaMalloc(&p, h.size() * sizeof(float)));
    net->allocs.push_back(p);
    CT_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    *d = (const float*)p;
    return 0;
  };
  const float* w = weights;
  const float* b = biases;
  const float* q = bn;
  for (int l = 0; l < 5; ++l) {
    const int in = dims[l], o = dims[l + 1];
    std::vector<float> wt((size_t)in * o), bb(b, b + o);
    for (int i = 0; i < o; ++i)
      for (int k = 0; k < in; ++k) wt[(size_t)k * o + i] = w[(size_t)i * in + k];   // torch [out][in] -> [in][out]
    int rc;
    if ((rc = up(wt, &net->dev.W[l])) || (rc = up(bb, &net->dev.b[l]))) return rc;
    w += (size_t)in * o;
    b += o;
    if (l < 4) {
      std::vector<float> sc(o), mu(o), be(o);
      for (int i = 0; i < o; ++i) {
        const float gamma = q[i], beta = q[o + i], mean = q[2 * o + i], var = q[3 * o + i];
        sc[i] = gamma / sqrtf(var + bn_eps);
        mu[i] = mean;
        be[i] = beta;
      }
      if ((rc = up(sc, &net->dev.bn_scale[l])) || (rc = up(mu, &net->dev.bn_mean[l])) || (rc = up(be, &net->dev.bn_beta[l]))) return rc;
      q += 4 * o;
    }
  }
  CT_CUDA(cudaFuncSetAttribute(chd_k_contact_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CT_TM * (1024 + 512) * sizeof(float))));
  *out = net;
  return 0;
}

void chd_contact_destroy(chd_contact_net* net) {
  if (!net) return;
  for (void* p : net->allocs) cudaFree(p);
  if (net->stream) cudaStreamDestroy(net->stream);
  delete net;
}

int chd_contact_forward_device(chd_contact_net* net, const double* frames_dev, int32_t V, int32_t Fmax, const int32_t* seq_lens_dev,
                               int64_t* labels_dev, float* logits_dev, float* min_abs_dev, void* stream) {
  if (!net || !frames_dev || Fmax < CT_WIN) return -1;
  cudaStream_t s = stream ? (cudaStream_t)stream : net->stream;
  const int Wn = Fmax - (CT_WIN - 1), total = V * Wn;
  const float big = 3.4e38f;
  CT_CUDA(cudaMemcpyAsync(min_abs_dev, &big, sizeof(float), cudaMemcpyHostToDevice, s));
  chd_k_contact_mlp<<<(total + CT_TM - 1) / CT_TM, CT_THREADS, CT_TM * (1024 + 512) * sizeof(float), s>>>(net->dev, frames_dev, V, Fmax, logits_dev);
  chd_k_contact_vote<<<(V * Fmax * 4 + 255) / 256, 256, 0, s>>>(logits_dev, V, Fmax, seq_lens_dev, (long long*)labels_dev, min_abs_dev);
  net->launches += 2;
  CT_CUDA(cudaGetLastError());
  return 0;
}

int chd_contact_forward(chd_contact_net* net, const double* frames, int32_t V, int32_t Fmax, const int32_t* seq_lens, int64_t* labels,
                        float* logits, float* min_abs_logit) {
  if (!net || !frames || !seq_lens || !labels || V <= 0 || Fmax < CT_WIN) return -1;
  const size_t nfr = (size_t)V * Fmax * 75, Wn = Fmax - (CT_WIN - 1), nlog = (size_t)V * Wn * 20, nlab = (size_t)V * Fmax * 4;
  double* d_fr = nullptr;
  int* d_len = nullptr;
  long long* d_lab = nullptr;
  float *d_log = nullptr, *d_min = nullptr;
  CT_CUDA(cudaMalloc((void**)&d_fr, nfr * sizeof(double)));
  CT_CUDA(cudaMalloc((void**)&d_len, V * sizeof(int)));
  CT_CUDA(cudaMalloc((void**)&d_lab, nlab * sizeof(long long)));
  CT_CUDA(cudaMalloc((void**)&d_log, nlog * sizeof(float)));
  CT_CUDA(cudaMalloc((void**)&d_min, sizeof(float)));
  CT_CUDA(cudaMemcpyAsync(d_fr, frames, nfr * sizeof(double), cudaMemcpyHostToDevice, net->stream));
  CT_CUDA(cudaMemcpyAsync(d_len, seq_lens, V * sizeof(int), cudaMemcpyHostToDevice, net->stream));
  int rc = chd_contact_forward_device(net, d_fr, V, Fmax, d_len, (int64_t*)d_lab, d_log, d_min, net->stream);
  if (rc) return rc;
  CT_CUDA(cudaMemcpyAsync(labels, d_lab, nlab * sizeof(long long), cudaMemcpyDeviceToHost, net->stream));
  if (logits) CT_CUDA(cudaMemcpyAsync(logits, d_log, nlog * sizeof(float), cudaMemcpyDeviceToHost, net->stream));
  if (min_abs_logit) CT_CUDA(cudaMemcpyAsync(min_abs_logit, d_min, sizeof(float), cudaMemcpyDeviceToHost, net->stream));
  CT_CUDA(cudaStreamSynchronize(net->stream));
  cudaFree(d_fr), cudaFree(d_len), cudaFree(d_lab), cudaFree(d_log), cudaFree(d_min);
  return 0;
}

int64_t chd_contact_launch_count(const chd_contact_net* net) { return net ? net->launches : 0; }

}  // extern "C"
