/* oracle/cpu_ref.c - plain-C restatement of SURVEY.md 8(a) Specs A-H.
 *
 * TEST INFRASTRUCTURE / timed CPU baseline only (bench.py cpu_baseline and
 * --impl reference, kind "port").  Never linked into the product library.
 * PARITY UNPINNED by the reference (/root/reference holds README.md + LICENSE
 * only); the contract is mxnet 1.3.0 contrib ROIAlign / box_nms /
 * MultiProposal (/root/reference/README.md:37) in the module roles of
 * /root/reference/README.md:16-17,24,28,32.  Checked against oracle/ *.py files
 * (tests/test_oracle.py: C port vs NumPy restatement), which in turn are checked against
 * torchvision CPU fixtures (tests/golden) and the independent routes of tests/test_cross_oracle.py.
 *
 * Parallelisation mirrors the reference stack: OpenMP over RoIs in RoIAlign
 * forward (as mxnet's roi_align.cc), serial backward (as mxnet's), OpenMP over
 * images elsewhere (the Python per-image loop of the reference would be serial;
 * threads are granted to keep the baseline generous).
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off -fPIC -shared  (see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int ora_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- Spec A ---- */
typedef struct { int lo, hi, valid; float l, h; } tap_t;

static tap_t axis_tap(float start, float bin, int grid, int p, int i, int size) {
  tap_t t;
  float c = (start + (float)p * bin) + (((float)i + 0.5f) * bin) / (float)grid;
  t.valid = !(c < -1.0f || c > (float)size);
  if (c <= 0.0f) c = 0.0f;
  int lo = t.valid ? (int)c : 0;
  if (lo >= size - 1) { t.hi = t.lo = size - 1; c = (float)t.lo; }
  else { t.lo = lo; t.hi = lo + 1; }
  t.l = c - (float)t.lo;
  t.h = 1.0f - t.l;
  return t;
}

typedef struct { int b, gh, gw; float rsw, rsh, bh, bw; } geom_t;

static geom_t roi_geom(const float* r, float scale, int PH, int PW, int sr) {
  geom_t g;
  g.b = (int)r[0];
  g.rsw = r[1] * scale; g.rsh = r[2] * scale;
  float rew = r[3] * scale, reh = r[4] * scale;
  float rw = fmaxf(rew - g.rsw, 1.0f), rh = fmaxf(reh - g.rsh, 1.0f);
  g.bh = rh / (float)PH; g.bw = rw / (float)PW;
  g.gh = sr > 0 ? sr : (int)ceilf(rh / (float)PH);
  g.gw = sr > 0 ? sr : (int)ceilf(rw / (float)PW);
  return g;
}

/* levels == NULL: single map (maps[0]).  maps[l] is (N,C,H[l],W[l]). */
void ora_roi_align_forward(const float* const* maps, const int* Hs, const int* Ws, const float* scales,
                           int num_levels, int N, int C, const float* rois, const int* levels, int R,
                           int PH, int PW, int sr, float* out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int n = 0; n < R; ++n) {
    int lv = (levels && num_levels > 1) ? levels[n] : 0;
    float* o = out + (size_t)n * C * PH * PW;
    geom_t g = roi_geom(rois + (size_t)n * 5, scales[lv < 0 || lv >= num_levels ? 0 : lv], PH, PW, sr);
    if (g.b < 0 || g.b >= N || lv < 0 || lv >= num_levels) { memset(o, 0, sizeof(float) * C * PH * PW); continue; }
    const int H = Hs[lv], W = Ws[lv];
    const float count = (float)(g.gh * g.gw);
    tap_t* ty = (tap_t*)malloc(sizeof(tap_t) * PH * g.gh);
    tap_t* tx = (tap_t*)malloc(sizeof(tap_t) * PW * g.gw);
    for (int t = 0; t < PH * g.gh; ++t) ty[t] = axis_tap(g.rsh, g.bh, g.gh, t / g.gh, t % g.gh, H);
    for (int t = 0; t < PW * g.gw; ++t) tx[t] = axis_tap(g.rsw, g.bw, g.gw, t / g.gw, t % g.gw, W);
    for (int c = 0; c < C; ++c) {
      const float* D = maps[lv] + ((size_t)g.b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.0f;
          for (int iy = 0; iy < g.gh; ++iy) {
            const tap_t y = ty[ph * g.gh + iy];
            if (!y.valid) continue;
            for (int ix = 0; ix < g.gw; ++ix) {
              const tap_t x = tx[pw * g.gw + ix];
              if (!x.valid) continue;
              const float w1 = y.h * x.h, w2 = y.h * x.l, w3 = y.l * x.h, w4 = y.l * x.l;
              acc += ((w1 * D[y.lo * W + x.lo] + w2 * D[y.lo * W + x.hi]) + w3 * D[y.hi * W + x.lo]) +
                     w4 * D[y.hi * W + x.hi];
            }
          }
          o[(c * PH + ph) * PW + pw] = acc / count;
        }
    }
    free(ty); free(tx);
  }
}

/* gmaps[l] must be zero-filled by the caller for req=write.  Serial, as mxnet 1.3 roi_align.cc. */
void ora_roi_align_backward(float* const* gmaps, const int* Hs, const int* Ws, const float* scales,
                            int num_levels, int N, int C, const float* rois, const int* levels, int R,
                            int PH, int PW, int sr, const float* gout) {
  for (int n = 0; n < R; ++n) {
    int lv = (levels && num_levels > 1) ? levels[n] : 0;
    if (lv < 0 || lv >= num_levels) continue;
    geom_t g = roi_geom(rois + (size_t)n * 5, scales[lv], PH, PW, sr);
    if (g.b < 0 || g.b >= N) continue;
    const int H = Hs[lv], W = Ws[lv];
    const float count = (float)(g.gh * g.gw);
    const float* go = gout + (size_t)n * C * PH * PW;
    for (int c = 0; c < C; ++c) {
      float* D = gmaps[lv] + ((size_t)g.b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          const float gv = go[(c * PH + ph) * PW + pw];
          for (int iy = 0; iy < g.gh; ++iy) {
            const tap_t y = axis_tap(g.rsh, g.bh, g.gh, ph, iy, H);
            if (!y.valid) continue;
            for (int ix = 0; ix < g.gw; ++ix) {
              const tap_t x = axis_tap(g.rsw, g.bw, g.gw, pw, ix, W);
              if (!x.valid) continue;
              D[y.lo * W + x.lo] += (gv * (y.h * x.h)) / count;
              D[y.lo * W + x.hi] += (gv * (y.h * x.l)) / count;
              D[y.hi * W + x.lo] += (gv * (y.l * x.h)) / count;
              D[y.hi * W + x.hi] += (gv * (y.l * x.l)) / count;
            }
          }
        }
    }
  }
}

/* ---------------------------------------------------------------- Spec G ---- */
void ora_map_roi_levels(const float* rois, int cols, int R, int L, float finest, int* out) {
  for (int i = 0; i < R; ++i) {
    const float* r = rois + (size_t)i * cols + (cols == 5 ? 1 : 0);
    float w = (r[2] - r[0]) + 1.0f, h = (r[3] - r[1]) + 1.0f;
    float v = sqrtf(w * h) / finest + 1e-6f;
    int lvl = 0; float t = 2.0f;
    for (int k = 1; k < L; ++k) { lvl += (v >= t); t *= 2.0f; }
    out[i] = lvl;
  }
}

/* ---------------------------------------------------------------- Spec B ---- */
typedef struct { float s; int i; } si_t;
static int cmp_si(const void* a, const void* b) {
  const si_t* x = (const si_t*)a; const si_t* y = (const si_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);   /* ties: lower index first => total order, qsort is safe */
}

static int sorted_order(const float* scores, int n, float valid_thresh, const unsigned char* valid, int topk,
                        si_t* ord) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (!(scores[i] > valid_thresh) && !(isinf(valid_thresh) && valid_thresh < 0 && scores[i] == scores[i])) continue;
    if (valid && !valid[i]) continue;   /* min-size filter (Spec H) drops rows before sorting */
    ord[m].s = scores[i] + 0.0f; ord[m].i = i; ++m;
  }
  qsort(ord, m, sizeof(si_t), cmp_si);
  if (topk > 0 && m > topk) m = topk;
  return m;
}

int ora_nms(const float* boxes, const float* scores, const int* ids, int n, float thr, float delta, int topk,
            float valid_thresh, int force_suppress, int max_out, const unsigned char* valid, int* keep) {
  si_t* ord = (si_t*)malloc(sizeof(si_t) * (n > 0 ? n : 1));
  int m = sorted_order(scores, n, valid_thresh, valid, topk, ord);
  float* area = (float*)malloc(sizeof(float) * (m > 0 ? m : 1));
  unsigned char* sup = (unsigned char*)calloc(m > 0 ? m : 1, 1);
  for (int r = 0; r < m; ++r) {
    const float* b = boxes + (size_t)ord[r].i * 4;
    area[r] = fmaxf((b[2] - b[0]) + delta, 0.0f) * fmaxf((b[3] - b[1]) + delta, 0.0f);
  }
  int nk = 0;
  for (int r = 0; r < m; ++r) {
    if (sup[r]) continue;
    keep[nk++] = ord[r].i;
    if (max_out > 0 && nk >= max_out) break;
    const float* a = boxes + (size_t)ord[r].i * 4;
    for (int p = r + 1; p < m; ++p) {
      if (sup[p]) continue;
      if (ids && !force_suppress && ids[ord[p].i] != ids[ord[r].i]) continue;
      const float* b = boxes + (size_t)ord[p].i * 4;
      const float iw = (fminf(a[2], b[2]) - fmaxf(a[0], b[0])) + delta;
      const float ih = (fminf(a[3], b[3]) - fmaxf(a[1], b[1])) + delta;
      if (iw <= 0.0f || ih <= 0.0f) continue;
      const float inter = iw * ih;
      if (inter / ((area[r] + area[p]) - inter) > thr) sup[p] = 1;
    }
  }
  free(ord); free(area); free(sup);
  return nk;
}

/* ------------------------------------------------------------ Specs D, E ---- */
static float iou_d(const float* a, const float* g, float d) {
  const float aa = ((a[2] - a[0]) + d) * ((a[3] - a[1]) + d), ag = ((g[2] - g[0]) + d) * ((g[3] - g[1]) + d);
  const float iw = (fminf(a[2], g[2]) - fmaxf(a[0], g[0])) + d, ih = (fminf(a[3], g[3]) - fmaxf(a[1], g[1])) + d;
  const float inter = (iw > 0.0f && ih > 0.0f) ? iw * ih : 0.0f;
  return inter / ((ag + aa) - inter);
}

void ora_bbox_overlaps(const float* b1, int G, const float* b2, int N, float delta, float* out) {
#pragma omp parallel for
  for (int g = 0; g < G; ++g)
    for (int n = 0; n < N; ++n) out[(size_t)g * N + n] = iou_d(b2 + (size_t)n * 4, b1 + (size_t)g * 4, delta);
}

/* One image.  Two sweeps over the implicit G x N matrix (no materialisation). */
void ora_max_iou_assign(const float* anchors, int N, const float* gts, int G, const int* gt_labels,
                        const unsigned char* flags, float pos, float neg, float min_pos, float delta,
                        int* assigned, float* max_ov, int* labels) {
  float* gt_max = (float*)malloc(sizeof(float) * (G > 0 ? G : 1));
  for (int g = 0; g < G; ++g) gt_max[g] = -INFINITY;
  for (int n = 0; n < N; ++n) {
    assigned[n] = -1; max_ov[n] = 0.0f; if (labels) labels[n] = 0;
    if (flags && !flags[n]) continue;
    if (G == 0) { assigned[n] = 0; continue; }
    float best = -INFINITY; int arg = 0;
    for (int g = 0; g < G; ++g) {
      const float v = iou_d(anchors + (size_t)n * 4, gts + (size_t)g * 4, delta);
      if (v > best) { best = v; arg = g; }
      if (v != v && best == best) { best = v; arg = g; }          /* NaN propagates (numpy / torch max), first NaN g */
      if (v > gt_max[g] || v != v) gt_max[g] = v;                 /* ... and through the per-GT maximum */
    }
    max_ov[n] = best;
    if (best >= 0.0f && best < neg) assigned[n] = 0;
    if (best >= pos) assigned[n] = arg + 1;
  }
  for (int n = 0; n < N && G > 0; ++n) {
    if (flags && !flags[n]) continue;
    for (int g = 0; g < G; ++g)
      if (gt_max[g] >= min_pos && iou_d(anchors + (size_t)n * 4, gts + (size_t)g * 4, delta) == gt_max[g])
        assigned[n] = g + 1;
    if (labels) labels[n] = (assigned[n] > 0 && gt_labels) ? gt_labels[assigned[n] - 1] : 0;
  }
  free(gt_max);
}

void ora_max_iou_assign_batch(const float* anchors, int N, const float* gts, const int* num_gts, int B, int Gmax,
                              const int* gt_labels, const unsigned char* flags, float pos, float neg,
                              float min_pos, float delta, int* assigned, float* max_ov, int* labels) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b)
    ora_max_iou_assign(anchors, N, gts + (size_t)b * Gmax * 4, num_gts ? num_gts[b] : Gmax,
                       gt_labels ? gt_labels + (size_t)b * Gmax : NULL, flags, pos, neg, min_pos, delta,
                       assigned + (size_t)b * N, max_ov + (size_t)b * N, labels ? labels + (size_t)b * N : NULL);
}

/* ---------------------------------------------------------------- Spec F ---- */
static float exp_cr(float x) { return (float)exp((double)x); }
static float log_cr(float x) { return (float)log((double)x); }

static void decode_one(const float* r, const float* dl, const float* means, const float* stds, float max_ratio,
                       int clip, float hmax, float wmax, float* o) {
  float dx = dl[0] * stds[0] + means[0], dy = dl[1] * stds[1] + means[1];
  float dw = dl[2] * stds[2] + means[2], dh = dl[3] * stds[3] + means[3];
  dw = fminf(fmaxf(dw, -max_ratio), max_ratio); dh = fminf(fmaxf(dh, -max_ratio), max_ratio);
  const float px = (r[0] + r[2]) * 0.5f, py = (r[1] + r[3]) * 0.5f;
  const float pw = (r[2] - r[0]) + 1.0f, ph = (r[3] - r[1]) + 1.0f;
  const float gw = pw * exp_cr(dw), gh = ph * exp_cr(dh);
  const float gx = px + pw * dx, gy = py + ph * dy;
  o[0] = (gx - gw * 0.5f) + 0.5f; o[1] = (gy - gh * 0.5f) + 0.5f;
  o[2] = (gx + gw * 0.5f) - 0.5f; o[3] = (gy + gh * 0.5f) - 0.5f;
  if (clip) {
    o[0] = fminf(fmaxf(o[0], 0.0f), wmax); o[2] = fminf(fmaxf(o[2], 0.0f), wmax);
    o[1] = fminf(fmaxf(o[1], 0.0f), hmax); o[3] = fminf(fmaxf(o[3], 0.0f), hmax);
  }
}

void ora_delta2bbox(const float* rois, const float* deltas, int m, const float* means, const float* stds,
                    int max_h, int max_w, double wh_ratio_clip, float* out) {
  const float max_ratio = (float)fabs(log(wh_ratio_clip));
  for (int i = 0; i < m; ++i)
    decode_one(rois + (size_t)i * 4, deltas + (size_t)i * 4, means, stds, max_ratio, max_h > 0 && max_w > 0,
               (float)(max_h - 1), (float)(max_w - 1), out + (size_t)i * 4);
}

void ora_bbox2delta(const float* p, const float* g, int m, const float* means, const float* stds, float* out) {
  for (int i = 0; i < m; ++i) {
    const float* a = p + (size_t)i * 4; const float* b = g + (size_t)i * 4;
    const float px = (a[0] + a[2]) * 0.5f, py = (a[1] + a[3]) * 0.5f, pw = (a[2] - a[0]) + 1.0f, ph = (a[3] - a[1]) + 1.0f;
    const float gx = (b[0] + b[2]) * 0.5f, gy = (b[1] + b[3]) * 0.5f, gw = (b[2] - b[0]) + 1.0f, gh = (b[3] - b[1]) + 1.0f;
    const float d[4] = {(gx - px) / pw, (gy - py) / ph, log_cr(gw / pw), log_cr(gh / ph)};
    for (int j = 0; j < 4; ++j) out[(size_t)i * 4 + j] = (d[j] - means[j]) / stds[j];
  }
}

/* ---------------------------------------------------------------- Spec H ---- */
typedef struct {
  int num_levels;
  int feat_h[8], feat_w[8];
  float stride[8];
  int num_base;
  float base_anchors[8][16][4];
  int nms_pre, nms_post, max_num;
  float nms_thr, min_bbox_size;
  float means[4], stds[4];
  float delta;
  double wh_ratio_clip;
} ora_rpn_config;

/* scores[l]: (B, n_l); deltas[l]: (B, n_l, 4); img_shapes (B,2); out (B,max_num,5); num_valid (B). */
void ora_rpn_proposals(const float* const* scores, const float* const* deltas, const int* img_shapes, int B,
                       const ora_rpn_config* c, float* out, int* num_valid) {
  const float max_ratio = (float)fabs(log(c->wh_ratio_clip));
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const int L = c->num_levels;
    int cap = 0;
    for (int l = 0; l < L; ++l) {
      int n = c->feat_h[l] * c->feat_w[l] * c->num_base;
      int k = (c->nms_pre > 0 && c->nms_pre < n) ? c->nms_pre : n;
      cap += (c->nms_post > 0 && c->nms_post < k) ? c->nms_post : k;
    }
    float* cat = (float*)malloc(sizeof(float) * 5 * (cap > 0 ? cap : 1));
    int total = 0;
    for (int l = 0; l < L; ++l) {
      const int W = c->feat_w[l], A = c->num_base;
      const int n = c->feat_h[l] * W * A;
      if (n == 0) continue;
      const float* s = scores[l] + (size_t)b * n;
      const float* d = deltas[l] + (size_t)b * n * 4;
      si_t* ord = (si_t*)malloc(sizeof(si_t) * n);
      const int k = sorted_order(s, n, -INFINITY, NULL, c->nms_pre, ord);
      float* boxes = (float*)malloc(sizeof(float) * 4 * k);
      float* sc = (float*)malloc(sizeof(float) * k);
      unsigned char* valid = (unsigned char*)malloc(k);
      for (int j = 0; j < k; ++j) {
        const int i = ord[j].i, a = i % A, cell = i / A, x = cell % W, y = cell / W;
        const float sx = (float)x * c->stride[l], sy = (float)y * c->stride[l];
        const float anc[4] = {c->base_anchors[l][a][0] + sx, c->base_anchors[l][a][1] + sy,
                              c->base_anchors[l][a][2] + sx, c->base_anchors[l][a][3] + sy};
        decode_one(anc, d + (size_t)i * 4, c->means, c->stds, max_ratio, 1, (float)(img_shapes[2 * b] - 1),
                   (float)(img_shapes[2 * b + 1] - 1), boxes + (size_t)j * 4);
        sc[j] = s[i];
        valid[j] = 1;
        if (c->min_bbox_size > 0.0f) {
          const float w = (boxes[j * 4 + 2] - boxes[j * 4]) + 1.0f, h = (boxes[j * 4 + 3] - boxes[j * 4 + 1]) + 1.0f;
          valid[j] = (w >= c->min_bbox_size && h >= c->min_bbox_size);
        }
      }
      int* keep = (int*)malloc(sizeof(int) * (k > 0 ? k : 1));
      const int nk = ora_nms(boxes, sc, NULL, k, c->nms_thr, c->delta, -1, -INFINITY, 1, c->nms_post, valid, keep);
      for (int j = 0; j < nk; ++j) {
        memcpy(cat + (size_t)total * 5, boxes + (size_t)keep[j] * 4, sizeof(float) * 4);
        cat[(size_t)total * 5 + 4] = sc[keep[j]];
        ++total;
      }
      free(ord); free(boxes); free(sc); free(valid); free(keep);
    }
    float* o = out + (size_t)b * c->max_num * 5;
    memset(o, 0, sizeof(float) * 5 * c->max_num);
    if (total > c->max_num) {
      si_t* ord = (si_t*)malloc(sizeof(si_t) * total);
      for (int j = 0; j < total; ++j) { ord[j].s = cat[(size_t)j * 5 + 4] + 0.0f; ord[j].i = j; }
      qsort(ord, total, sizeof(si_t), cmp_si);
      for (int j = 0; j < c->max_num; ++j) memcpy(o + (size_t)j * 5, cat + (size_t)ord[j].i * 5, sizeof(float) * 5);
      free(ord);
      total = c->max_num;
    } else {
      memcpy(o, cat, sizeof(float) * 5 * total);
    }
    num_valid[b] = total;
    free(cat);
  }
}

int ora_sizeof_rpn_config(void) { return (int)sizeof(ora_rpn_config); }
