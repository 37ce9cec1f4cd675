/*
 * ORACLE (plain C) — test infrastructure, not product code.
 *
 * An ATen-independent restatement of the arithmetic on the logits -> loss -> metrics path, in double precision,
 * used by tests/ as a second checker next to oracle/oracle.py (which re-issues the reference's ATen calls).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Follows (reference file:line):
 *   oc_resize_bilinear  utils/ops.py:26 -> F.interpolate(mode='bilinear'); index rule
 *                       torch/include/ATen/native/UpSample.h:271-312,442-476 (fp32 index math, as ATen)
 *   oc_ce               models/losses/cross_entropy_loss.py:56-72 + models/losses/utils.py:60-80
 *   oc_dice             models/losses/dice_loss.py:31-58,117-133
 *   oc_accuracy_top1    models/losses/accuracy.py:41-60
 *   oc_argmax           core/evaluation/metrics.py:106 (argmax of logits; soft-max is monotone up to rounding)
 *   oc_areas            core/evaluation/metrics.py:236-270
 *   oc_lovasz           models/losses/lovasz_loss.py:26-231 (lovasz_grad, lovasz_softmax(_flat), lovasz_hinge(_flat)) +
 *                       the reductions of LovaszLoss.forward :272-298; the Jaccard increments in exact (integer-count) form
 *
 * Parity pinning: checked against tests/golden/ (recorded from the reference's own files) in tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float scale_of(int in, int out, int ac) {
  if (ac) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}

static void src_index(float scale, int dst, int in, int ac, int* i0, int* i1, float* l1) {
  float src;
  if (ac) {
    src = scale * (float)dst;
  } else {
    src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  int i = (int)src;
  if (i > in - 1) i = in - 1;
  *i0 = i;
  *i1 = i + (i < in - 1 ? 1 : 0);
  float l = src - (float)i;
  if (l < 0.f) l = 0.f;
  if (l > 1.f) l = 1.f;
  *l1 = l;
}

/* out (NC,H,W) <- in (NC,h,w); double accumulation of the four taps. */
void oc_resize_bilinear(const float* in, double* out, int NC, int h, int w, int H, int W, int ac) {
  const float sh = scale_of(h, H, ac), sw = scale_of(w, W, ac);
#pragma omp parallel for
  for (int nc = 0; nc < NC; ++nc) {
    for (int Y = 0; Y < H; ++Y) {
      int y0, y1;
      float ly;
      src_index(sh, Y, h, ac, &y0, &y1, &ly);
      for (int X = 0; X < W; ++X) {
        int x0, x1;
        float lx;
        src_index(sw, X, w, ac, &x0, &x1, &lx);
        const float* p = in + (size_t)nc * h * w;
        const double h1 = ly, h0 = 1.0 - (double)ly, w1 = lx, w0 = 1.0 - (double)lx;
        out[((size_t)nc * H + Y) * W + X] = h0 * (w0 * p[y0 * w + x0] + w1 * p[y0 * w + x1]) +
                                            h1 * (w0 * p[y1 * w + x0] + w1 * p[y1 * w + x1]);
      }
    }
  }
}

/* grad_in (NC,h,w) += transpose of the above applied to grad_out (NC,H,W). */
static void resize_bilinear_bwd(const double* go, double* gi, int NC, int h, int w, int H, int W, int ac) {
  const float sh = scale_of(h, H, ac), sw = scale_of(w, W, ac);
#pragma omp parallel for
  for (int nc = 0; nc < NC; ++nc) {
    double* g = gi + (size_t)nc * h * w;
    for (int Y = 0; Y < H; ++Y) {
      int y0, y1;
      float ly;
      src_index(sh, Y, h, ac, &y0, &y1, &ly);
      for (int X = 0; X < W; ++X) {
        int x0, x1;
        float lx;
        src_index(sw, X, w, ac, &x0, &x1, &lx);
        const double v = go[((size_t)nc * H + Y) * W + X];
        const double h1 = ly, h0 = 1.0 - (double)ly, w1 = lx, w0 = 1.0 - (double)lx;
        g[y0 * w + x0] += h0 * w0 * v;
        g[y0 * w + x1] += h0 * w1 * v;
        g[y1 * w + x0] += h1 * w0 * v;
        g[y1 * w + x1] += h1 * w1 * v;
      }
    }
  }
}

/*
 * Cross-entropy (+ top-1 accuracy) on logits (N,C,h,w) resized to (H,W).
 *   reduction: 0 none, 1 mean, 2 sum; avg_factor < 0 means "not given".
 * Outputs: loss_out[0] (scalar reductions) or loss_px (N,H,W) for 'none'; grad (N,C,h,w) for upstream gradient 1
 * (or grad_px per pixel for 'none'); counts[0..2] = n_valid, n_correct, n_acc.
 */
void oc_ce(const float* logits, const int64_t* labels, const float* pixel_weight, const float* class_weight, int N, int C,
           int h, int w, int H, int W, int ac, int64_t ignore_index, int reduction, int avg_non_ignore, double avg_factor,
           double loss_weight, int acc_has_ignore, int64_t acc_ignore, const double* grad_px, double* loss_out,
           double* loss_px, double* grad, int64_t* counts) {
  const size_t HW = (size_t)H * W, hw = (size_t)h * w;
  double* full = (double*)malloc(sizeof(double) * (size_t)N * C * HW);
  double* gfull = (double*)calloc((size_t)N * C * HW, sizeof(double));
  oc_resize_bilinear(logits, full, N * C, h, w, H, W, ac);
  const double eps = 1.1920928955078125e-07;
  double total = 0.0;
  int64_t n_valid = 0, n_correct = 0, n_acc = 0;
  for (int n = 0; n < N; ++n)
    for (size_t p = 0; p < HW; ++p) {
      const int64_t y = labels[(size_t)n * HW + p];
      n_valid += (y != ignore_index);
    }
  double denom = 1.0;
  if (reduction == 1) {
    if (avg_factor >= 0.0) denom = (double)(float)(avg_factor + eps);
    else if (avg_non_ignore) denom = (double)(float)((double)n_valid + eps);
    else denom = (double)N * (double)HW;
  }
  for (int n = 0; n < N; ++n) {
    for (size_t p = 0; p < HW; ++p) {
      const double* z = full + (size_t)n * C * HW + p;
      double m = -INFINITY;
      int arg = 0;
      for (int c = 0; c < C; ++c)
        if (z[c * HW] > m) { m = z[c * HW]; arg = c; }
      double s = 0.0;
      for (int c = 0; c < C; ++c) s += exp(z[c * HW] - m);
      const double lse = m + log(s);
      const int64_t y = labels[(size_t)n * HW + p];
      const int av = acc_has_ignore ? (y != acc_ignore) : 1;
      n_acc += av;
      n_correct += (av && (int64_t)arg == y);
      double l = 0.0;
      if (y != ignore_index && y >= 0 && y < C) {
        const double wt = (class_weight ? class_weight[y] : 1.0) * (pixel_weight ? pixel_weight[(size_t)n * HW + p] : 1.0);
        l = wt * (lse - z[y * HW]);
        const double up = (grad_px ? grad_px[(size_t)n * HW + p] : 1.0) * loss_weight / denom;
        for (int c = 0; c < C; ++c) {
          const double pr = exp(z[c * HW] - lse);
          gfull[((size_t)n * C + c) * HW + p] = up * wt * (pr - (c == y ? 1.0 : 0.0));
        }
      }
      if (loss_px) loss_px[(size_t)n * HW + p] = loss_weight * l;
      total += l;
    }
  }
  if (loss_out) loss_out[0] = loss_weight * total / denom;
  if (grad) {
    memset(grad, 0, sizeof(double) * (size_t)N * C * hw);
    if (h == H && w == W) memcpy(grad, gfull, sizeof(double) * (size_t)N * C * hw);
    else resize_bilinear_bwd(gfull, grad, N * C, h, w, H, W, ac);
  }
  if (counts) { counts[0] = n_valid; counts[1] = n_correct; counts[2] = n_acc; }
  free(full);
  free(gfull);
}

/* Dice on logits (N,C,H,W) at label resolution; grad for upstream gradient 1. avg_factor < 0 = not given. */
void oc_dice(const float* logits, const int64_t* labels, const float* class_weight, int N, int C, int H, int W,
             int64_t ignore_index, double smooth, double exponent, double loss_weight, int reduction, double avg_factor,
             double* loss_out, double* grad) {
  const size_t HW = (size_t)H * W;
  const double eps = 1.1920928955078125e-07;
  double* prob = (double*)malloc(sizeof(double) * (size_t)N * C * HW);
  double* num = (double*)calloc((size_t)N * C, sizeof(double));
  double* den = (double*)calloc((size_t)N * C, sizeof(double));
  for (int n = 0; n < N; ++n)
    for (size_t p = 0; p < HW; ++p) {
      const float* z = logits + (size_t)n * C * HW + p;
      double m = -INFINITY, s = 0.0;
      for (int c = 0; c < C; ++c)
        if (z[c * HW] > m) m = z[c * HW];
      for (int c = 0; c < C; ++c) s += exp((double)z[c * HW] - m);
      int64_t y = labels[(size_t)n * HW + p];
      const double v = (y != ignore_index) ? 1.0 : 0.0;
      const int64_t yc = y < 0 ? 0 : (y > C - 1 ? C - 1 : y);
      for (int c = 0; c < C; ++c) {
        const double pr = exp((double)z[c * HW] - m) / s;
        prob[((size_t)n * C + c) * HW + p] = pr;
        const double t = (c == yc) ? 1.0 : 0.0;
        num[n * C + c] += pr * t * v;
        den[n * C + c] += pow(pr, exponent) + t;
      }
    }
  double K = loss_weight / ((double)C * (double)N);
  if (reduction == 1 && avg_factor >= 0.0) K /= (double)(float)(avg_factor + eps);
  double total = 0.0;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      if ((int64_t)c == ignore_index) continue;
      const double cw = class_weight ? class_weight[c] : 1.0;
      total += cw * (1.0 - (2.0 * num[n * C + c] + smooth) / (den[n * C + c] + smooth));
    }
  if (loss_out) loss_out[0] = K * total;
  if (grad) {
    for (int n = 0; n < N; ++n)
      for (size_t p = 0; p < HW; ++p) {
        int64_t y = labels[(size_t)n * HW + p];
        const double v = (y != ignore_index) ? 1.0 : 0.0;
        const int64_t yc = y < 0 ? 0 : (y > C - 1 ? C - 1 : y);
        double dot = 0.0;
        double g[4096];
        for (int c = 0; c < C; ++c) {
          const double pr = prob[((size_t)n * C + c) * HW + p];
          const double cw = ((int64_t)c == ignore_index) ? 0.0 : (class_weight ? class_weight[c] : 1.0);
          const double nm = 2.0 * num[n * C + c] + smooth, dn = den[n * C + c] + smooth;
          const double t = (c == yc) ? 1.0 : 0.0;
          g[c] = K * cw * (-2.0 * t * v / dn + nm / (dn * dn) * exponent * pow(pr, exponent - 1.0));
          dot += pr * g[c];
        }
        for (int c = 0; c < C; ++c) {
          const double pr = prob[((size_t)n * C + c) * HW + p];
          grad[((size_t)n * C + c) * HW + p] = pr * (g[c] - dot);
        }
      }
  }
  free(prob);
  free(num);
  free(den);
}

/* arg-max over classes of logits (C,HW) -> int64 (HW); lowest index wins ties. */
void oc_argmax(const float* logits, int C, int64_t HW, int64_t* out) {
#pragma omp parallel for
  for (int64_t p = 0; p < HW; ++p) {
    float m = -INFINITY;
    int arg = 0;
    for (int c = 0; c < C; ++c)
      if (logits[(size_t)c * HW + p] > m) { m = logits[(size_t)c * HW + p]; arg = c; }
    out[p] = arg;
  }
}

/* areas[0..C) intersect, [C..2C) pred, [2C..3C) label for one image; values outside [0,C-1] are dropped (histc). */
void oc_areas(const int64_t* pred, const float* gt, int64_t n, int C, int64_t ignore_index, int64_t* areas) {
  memset(areas, 0, sizeof(int64_t) * 3 * (size_t)C);
  for (int64_t i = 0; i < n; ++i) {
    const float g = gt[i];
    if (g == (float)ignore_index) continue;
    const int64_t p = pred[i];
    const int64_t gi = (int64_t)g;
    const int pin = (p >= 0 && p <= C - 1), gin = (g >= 0.f && g <= (float)(C - 1));
    if ((float)p == g && pin) areas[p] += 1;
    if (pin) areas[C + p] += 1;
    if (gin) areas[2 * C + gi] += 1;
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Lovasz-Softmax / Lovasz hinge (models/losses/lovasz_loss.py). One SEGMENT = the valid pixels of one image group
 * (the whole batch, or one image when per_image) for one class: errors sorted descending (ties by pixel index), the
 * Jaccard index of the prefix sets first-differenced (lovasz_grad :26-39, here from integer counts: 1/U for a foreground
 * item, I/(U(U-1)) otherwise, J_0 for the first), loss = sum e_i g_i.
 *   logits (N,C,HW) multi-class / (N,HW) binary (C == 1); labels (N,HW) int64
 *   class_mask (C) 1 = class takes part ('all' / list) ; only_present: skip classes without a foreground pixel (:153-154)
 *   loss_out: n_groups doubles when per_image && reduction == none (0), else 1 ; grad_out_w: upstream gradient per
 *   output (NULL = 1) ; grad (N,C,HW) doubles.                                                                      */
typedef struct { double e; int64_t idx; int fg; } lv_item;
static int lv_cmp(const void* a, const void* b) {
  const lv_item* x = (const lv_item*)a; const lv_item* y = (const lv_item*)b;
  if (x->e > y->e) return -1;
  if (x->e < y->e) return 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

void oc_lovasz(const float* logits, const int64_t* labels, const float* class_weight, const uint8_t* class_mask, int N, int C,
               int64_t HW, int has_ignore, int64_t ignore_index, int binary, int per_image, int only_present, int reduction,
               double avg_factor, double loss_weight, const double* grad_out_w, double* loss_out, double* grad) {
  const int n_groups = per_image ? N : 1;
  const int imgs = per_image ? 1 : N;
  const int64_t P = (int64_t)imgs * HW;
  const double eps32 = 1.1920928955078125e-07;
  double* prob = (double*)malloc(sizeof(double) * (size_t)N * C * HW);      /* soft-max probabilities (multi-class) */
  double* G = (double*)calloc((size_t)N * C * HW, sizeof(double));          /* d loss_out / d p   (binary: / d z)   */
  lv_item* items = (lv_item*)malloc(sizeof(lv_item) * (size_t)P);
  double* gl = (double*)malloc(sizeof(double) * (size_t)n_groups);
  memset(grad, 0, sizeof(double) * (size_t)N * C * HW);
  if (!binary) {
    for (int n = 0; n < N; ++n)
      for (int64_t i = 0; i < HW; ++i) {
        double m = -INFINITY, sum = 0.0;
        for (int c = 0; c < C; ++c) { const double z = logits[((size_t)n * C + c) * HW + i]; if (z > m) m = z; }
        for (int c = 0; c < C; ++c) sum += exp((double)logits[((size_t)n * C + c) * HW + i] - m);
        for (int c = 0; c < C; ++c) prob[((size_t)n * C + c) * HW + i] = exp((double)logits[((size_t)n * C + c) * HW + i] - m) / sum;
      }
  }
  double gscale = 1.0;
  if (per_image && reduction == 1) gscale = avg_factor >= 0.0 ? 1.0 / (double)(float)((float)avg_factor + (float)eps32) : 1.0 / n_groups;
  double total = 0.0;
  for (int g = 0; g < n_groups; ++g) {
    const int n0 = per_image ? g : 0;
    double* lc = (double*)calloc((size_t)C, sizeof(double));
    int* used = (int*)calloc((size_t)C, sizeof(int));
    int cnt = 0;
    for (int c = 0; c < C; ++c) {
      if (!binary && class_mask && !class_mask[c]) continue;
      int64_t m = 0, gts = 0;
      for (int nl = 0; nl < imgs; ++nl)
        for (int64_t i = 0; i < HW; ++i) {
          const int64_t y = labels[(size_t)(n0 + nl) * HW + i];
          if (has_ignore && y == ignore_index) continue;
          lv_item it;
          it.idx = (int64_t)nl * HW + i;
          if (binary) {
            it.fg = y != 0;
            it.e = 1.0 - (double)logits[(size_t)(n0 + nl) * HW + i] * (it.fg ? 1.0 : -1.0);
          } else {
            it.fg = (y == c);
            it.e = fabs((it.fg ? 1.0 : 0.0) - prob[((size_t)(n0 + nl) * C + c) * HW + i]);
          }
          gts += it.fg;
          items[m++] = it;
        }
      if (!binary && only_present && gts == 0) continue;
      if (m == 0 && !binary) { used[c] = 1; ++cnt; continue; }
      qsort(items, (size_t)m, sizeof(lv_item), lv_cmp);
      int64_t cum = 0;
      double loss = 0.0;
      for (int64_t i = 0; i < m; ++i) {
        cum += items[i].fg;
        const double I = (double)(gts - cum), U = (double)(gts + (i + 1) - cum);
        const double gi = i == 0 ? 1.0 - I / U : (items[i].fg ? 1.0 / U : I / (U * (U - 1.0)));
        const int nl = (int)(items[i].idx / HW);
        const int64_t px = items[i].idx - (int64_t)nl * HW;
        double* Gp = G + ((size_t)(n0 + nl) * C + c) * HW + px;
        if (binary) {
          const double e = items[i].e;
          if (e > 0.0) { loss += e * gi; *Gp = items[i].fg ? -gi : gi; }   /* d relu(1 - z*sign)/dz = -sign */
        } else {
          loss += items[i].e * gi;
          *Gp = items[i].fg ? -gi : gi;                                     /* d|fg - p|/dp */
        }
      }
      lc[c] = loss;
      used[c] = 1;
      ++cnt;
    }
    double sum = 0.0;
    for (int c = 0; c < C; ++c)
      if (used[c]) sum += ((class_weight && !binary) ? (double)class_weight[c] : 1.0) * lc[c];
    gl[g] = cnt ? sum / cnt : 0.0;
    /* coefficient of every class of this group inside the returned value(s), then the soft-max Jacobian */
    const int none_vec = per_image && reduction == 0;
    const double up = (grad_out_w ? grad_out_w[none_vec ? g : 0] : 1.0) * loss_weight * gscale;
    for (int nl = 0; nl < imgs; ++nl)
      for (int64_t i = 0; i < HW; ++i) {
        const size_t base = (size_t)(n0 + nl) * C * HW + i;
        if (has_ignore && labels[(size_t)(n0 + nl) * HW + i] == ignore_index) continue;
        if (binary) { grad[base] = up * G[base]; continue; }
        double dot = 0.0;
        for (int c = 0; c < C; ++c)
          if (used[c]) dot += ((class_weight ? (double)class_weight[c] : 1.0) / cnt) * G[base + (size_t)c * HW] * prob[base + (size_t)c * HW];
        for (int c = 0; c < C; ++c) {
          const double a = used[c] ? ((class_weight ? (double)class_weight[c] : 1.0) / cnt) * G[base + (size_t)c * HW] : 0.0;
          grad[base + (size_t)c * HW] = up * prob[base + (size_t)c * HW] * (a - dot);
        }
      }
    if (none_vec) loss_out[g] = loss_weight * gl[g];
    else total += gl[g];
    free(lc);
    free(used);
  }
  if (!(per_image && reduction == 0)) loss_out[0] = loss_weight * gscale * total;
  free(prob); free(G); free(items); free(gl);
}
