// Random positive / negative sampling and target packing.
// Module roles: mxdetection/core/anchor + core/bbox (/root/reference/README.md:16-17); RandomSampler and
// anchor_target_single / bbox_target_single of mmdet 0.5.
//
// Device RNG contract (the reference draws with numpy.random.choice, which no device can replay): every candidate
// carries a key u in [0,1); the sample is the candidates with the LARGEST keys, ties to the lower index.  That is
// the library's stable top-k, so sampling is three launches - mask the keys per class, one two-segment top-k,
// one finalising CTA - and the packed label / weight / target tensors are four memsets plus one scatter kernel.
// Nothing synchronises with the host: fixed-capacity index lists (-1 padded) and device counts.
#include "internal.h"

namespace mxd {

struct F4t { float v[4]; };

__global__ void sample_mask_keys_kernel(const int* __restrict__ assigned, const float* __restrict__ keys, int n,
                                        float* __restrict__ kpos, float* __restrict__ kneg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = assigned[i];
  const float k = keys[i];
  kpos[i] = a > 0 ? k : -1.0f;
  kneg[i] = a == 0 ? k : -1.0f;
}

// idx (2, kmax) / cnt (2): rows of the two-segment top-k (positives, negatives)
__global__ void sample_finalize_kernel(const int* __restrict__ idx, const int* __restrict__ cnt, int kmax, int kp,
                                       int kn, int num, int neg_pos_ub, int* __restrict__ pos_inds,
                                       int* __restrict__ neg_inds, int* __restrict__ counts) {
  const int num_pos = kp > 0 ? min(cnt[0], kp) : 0;
  int quota = num - num_pos;
  if (neg_pos_ub >= 0) quota = min(quota, neg_pos_ub * max(num_pos, 1));
  const int num_neg = kn > 0 ? max(min(cnt[1], min(quota, kn)), 0) : 0;
  for (int j = threadIdx.x; j < kp; j += blockDim.x) pos_inds[j] = j < num_pos ? idx[j] : -1;
  for (int j = threadIdx.x; j < kn; j += blockDim.x) neg_inds[j] = j < num_neg ? idx[kmax + j] : -1;
  if (threadIdx.x == 0) { counts[0] = num_pos; counts[1] = num_neg; }
}

__global__ void pack_targets_kernel(const float4* __restrict__ anchors, const int* __restrict__ assigned,
                                    const float4* __restrict__ gts, const int* __restrict__ gt_labels,
                                    const int* __restrict__ pos_inds, int kp, const int* __restrict__ neg_inds, int kn,
                                    F4t means, F4t stds, float pos_w, int* __restrict__ labels,
                                    float* __restrict__ label_w, float4* __restrict__ tgt, float4* __restrict__ tgt_w) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < kp) {
    const int i = pos_inds[j];
    if (i < 0) return;
    const int g = max(assigned[i] - 1, 0);
    tgt[i] = encode_box(anchors[i], gts[g], means.v, stds.v);
    tgt_w[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    labels[i] = gt_labels ? gt_labels[g] : 1;
    label_w[i] = pos_w;
  } else if (j < kp + kn) {
    const int i = neg_inds[j - kp];
    if (i >= 0) label_w[i] = 1.0f;
  }
}

}  // namespace mxd

using namespace mxd;

extern "C" {

static size_t sample_ws_layout(long long n, int kmax, size_t* off_idx, size_t* off_cnt) {
  size_t o = align_up((size_t)(2 * n) * sizeof(float), 256);
  *off_idx = o; o += align_up((size_t)2 * (kmax > 0 ? kmax : 1) * sizeof(int), 256);
  *off_cnt = o; o += 256;
  return o;
}

size_t mxd_random_sample_workspace_bytes(long long n, int num) {
  size_t a, b;
  return sample_ws_layout(n, num > 0 ? num : 1, &a, &b);
}

int mxd_random_sample(const DLTensor* assigned, const DLTensor* keys, int num, double pos_fraction, int neg_pos_ub,
                      DLTensor* pos_inds, DLTensor* neg_inds, DLTensor* counts, void* workspace,
                      size_t workspace_bytes, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(assigned, "assigned", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(keys, "keys", F32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(pos_inds, "pos_inds", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(neg_inds, "neg_inds", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(counts, "counts", I32, 1, 1, &dev))) return rc;
  const long long n = assigned->shape[0];
  MXD_REQUIRE(keys->shape[0] == n && n < (1ll << 31), MXD_EINVAL, "keys must be (N)");
  MXD_REQUIRE(num >= 0 && num <= MXD_SORT_CAP, MXD_ENOTSUP, "sample size %d exceeds the sort capacity %d", num, MXD_SORT_CAP);
  const int kp = (int)std::min<long long>((long long)((double)num * pos_fraction), n);
  const int kn = (int)std::min<long long>(num, n);
  MXD_REQUIRE(pos_inds->shape[0] == kp && neg_inds->shape[0] == kn && counts->shape[0] == 2, MXD_EINVAL,
              "pos_inds / neg_inds / counts must be (%d) / (%d) / (2)", kp, kn);
  cudaStream_t st = as_stream(stream);
  const int kmax = std::max(std::max(kp, kn), 1);
  size_t off_idx, off_cnt;
  const size_t need = sample_ws_layout(n, kmax, &off_idx, &off_cnt);
  MXD_REQUIRE(workspace && workspace_bytes >= need, MXD_EINVAL, "workspace too small: %zu < %zu", workspace_bytes, need);
  MXD_REQUIRE(((uintptr_t)workspace & 255) == 0, MXD_EINVAL, "workspace must be 256-byte aligned");
  char* base = static_cast<char*>(workspace);
  float* kpos = reinterpret_cast<float*>(base);
  float* kneg = kpos + n;
  int* idx = reinterpret_cast<int*>(base + off_idx);
  int* cnt = reinterpret_cast<int*>(base + off_cnt);
  if (n > 0) {
    sample_mask_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dptr<int>(assigned), dptr<float>(keys), (int)n, kpos, kneg);
    MXD_POST_LAUNCH("sample_mask_keys");
  }
  TopkParams p = {};
  p.num_levels = 2; p.batch = 1;
  p.scores[0] = kpos; p.scores[1] = kneg; p.seg_stride[0] = p.seg_stride[1] = 0; p.elem_stride = 1;
  p.n[0] = kp > 0 ? (int)n : 0; p.n[1] = kn > 0 ? (int)n : 0;     // an unused segment is an empty one
  p.k[0] = kp; p.k[1] = kn; p.kmax = kmax;
  p.valid_thresh = -0.5f;                                          // masked rows (-1) are dropped, keys >= 0 stay
  p.out_idx = idx; p.out_cnt = cnt;
  if ((rc = launch_topk(p, st))) return rc;
  sample_finalize_kernel<<<1, 256, 0, st>>>(idx, cnt, kmax, kp, kn, num, neg_pos_ub, dptr<int>(pos_inds), dptr<int>(neg_inds),
                                            dptr<int>(counts));
  MXD_POST_LAUNCH("sample_finalize");
  return MXD_OK;
}

int mxd_pack_targets(const DLTensor* anchors, const DLTensor* assigned, const DLTensor* gts, const DLTensor* gt_labels,
                     const DLTensor* pos_inds, const DLTensor* neg_inds, const float* means, const float* stds,
                     float pos_weight, DLTensor* labels, DLTensor* label_weights, DLTensor* bbox_targets,
                     DLTensor* bbox_weights, void* stream) {
  int dev = -1, rc;
  if ((rc = check_tensor(anchors, "anchors", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(assigned, "assigned", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(gts, "gts", F32, 2, 2, &dev))) return rc;
  if (gt_labels && (rc = check_tensor(gt_labels, "gt_labels", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(pos_inds, "pos_inds", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(neg_inds, "neg_inds", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(labels, "labels", I32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(label_weights, "label_weights", F32, 1, 1, &dev))) return rc;
  if ((rc = check_tensor(bbox_targets, "bbox_targets", F32, 2, 2, &dev))) return rc;
  if ((rc = check_tensor(bbox_weights, "bbox_weights", F32, 2, 2, &dev))) return rc;
  const long long n = anchors->shape[0], G = gts->shape[0];
  MXD_REQUIRE(anchors->shape[1] == 4 && gts->shape[1] == 4 && assigned->shape[0] == n, MXD_EINVAL, "anchors (N,4), gts (G,4), assigned (N)");
  MXD_REQUIRE(!gt_labels || gt_labels->shape[0] == G, MXD_EINVAL, "gt_labels must be (G)");
  MXD_REQUIRE(labels->shape[0] == n && label_weights->shape[0] == n && bbox_targets->shape[0] == n && bbox_targets->shape[1] == 4 &&
                  bbox_weights->shape[0] == n && bbox_weights->shape[1] == 4, MXD_EINVAL, "outputs must be (N) / (N) / (N,4) / (N,4)");
  MXD_REQUIRE(means && stds, MXD_EINVAL, "means/stds must be float[4]");
  MXD_REQUIRE((((uintptr_t)dptr<float>(anchors) | (uintptr_t)dptr<float>(gts) | (uintptr_t)dptr<float>(bbox_targets) |
                (uintptr_t)dptr<float>(bbox_weights)) & 15) == 0, MXD_EINVAL, "box tensors must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (n > 0) {
    MXD_CUDA_OK(cudaMemsetAsync(dptr<int>(labels), 0, (size_t)n * 4, st));
    MXD_CUDA_OK(cudaMemsetAsync(dptr<float>(label_weights), 0, (size_t)n * 4, st));
    MXD_CUDA_OK(cudaMemsetAsync(dptr<float>(bbox_targets), 0, (size_t)n * 16, st));
    MXD_CUDA_OK(cudaMemsetAsync(dptr<float>(bbox_weights), 0, (size_t)n * 16, st));
  }
  const int kp = (int)pos_inds->shape[0], kn = (int)neg_inds->shape[0];
  if (n == 0 || kp + kn == 0) return MXD_OK;
  MXD_REQUIRE(G > 0 || kp == 0, MXD_EINVAL, "positives without ground truth");
  F4t mu, sd;
  for (int j = 0; j < 4; ++j) { mu.v[j] = means[j]; sd.v[j] = stds[j]; }
  pack_targets_kernel<<<(kp + kn + 255) / 256, 256, 0, st>>>(
      reinterpret_cast<const float4*>(dptr<float>(anchors)), dptr<int>(assigned),
      reinterpret_cast<const float4*>(dptr<float>(gts)), gt_labels ? dptr<int>(gt_labels) : nullptr, dptr<int>(pos_inds), kp,
      dptr<int>(neg_inds), kn, mu, sd, pos_weight > 0.0f ? pos_weight : 1.0f, dptr<int>(labels), dptr<float>(label_weights),
      reinterpret_cast<float4*>(dptr<float>(bbox_targets)), reinterpret_cast<float4*>(dptr<float>(bbox_weights)));
  MXD_POST_LAUNCH("pack_targets");
  return MXD_OK;
}

}  // extern "C"
