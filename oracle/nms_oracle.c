/* CPU oracle for the greedy 3-D NMS of LSSD3D.detect_objects -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates /root/reference/lesions3d/ssd3d.py:407-426 (suppress loop over score-sorted boxes) with the IoU of
 * utils.py:105-149 (find_intersection3d / find_jaccard_overlap3d), one separately rounded fp32 operation per
 * reference operation (compile with -ffp-contract=off: no FMA contraction), for candidate lists the reference's
 * n x n matrix cannot hold (1 M .. 2.5 M boxes = 4 .. 25 TB).  Only tests/ may load it.
 *
 *   keep[i] = 1  <=>  no kept j < i has IoU(j, i) > thr          (strict >, NaN compares false)
 *
 * which is what the reference's loop computes: box i is skipped when an earlier KEPT box marked it
 * (ssd3d.py:417-418), otherwise it marks every box it overlaps (ssd3d.py:422) and stays (ssd3d.py:426).
 *
 * To make millions of boxes tractable on one core the kept boxes are binned twice -- by volume octave
 * (IoU <= min(va,vb)/max(va,vb), so for thr > 0 only boxes with vb in (thr*va, va/thr) can matter; applied with
 * a 1e-4 relative safety margin, and switched off for thr <= 0) and, per octave, in a uniform grid over the
 * minimum corner (a box k can only intersect q if q.min - maxext_level < k.min < q.max on every axis).  Both are
 * pure pruning: every surviving pair goes through the exact arithmetic above.  Pinned against the n x n
 * restatement (oracle/ssd3d_oracle.py greedy_nms) and the reference's own loop in tests/test_oracle_golden.py and
 * tests/test_oracle_vs_reference.py.
 *
 * Precondition (checked; returns -2 otherwise): thr >= 0 and every coordinate finite with max >= min -- the only
 * inputs for which "no intersection => cannot suppress" holds.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LEVELS 96
#define MAX_DIM 256

typedef struct {
  int used;
  float cell, maxext;
  float lo[3];
  int dim[3];
  int32_t* head; /* dim0*dim1*dim2 cell heads, -1 = empty */
} level_t;

static inline float iou_ref(const float* a, float va, const float* b, float vb) {
  /* utils.py:119-122 */
  float lo0 = a[0] > b[0] ? a[0] : b[0], lo1 = a[1] > b[1] ? a[1] : b[1], lo2 = a[2] > b[2] ? a[2] : b[2];
  float hi0 = a[3] < b[3] ? a[3] : b[3], hi1 = a[4] < b[4] ? a[4] : b[4], hi2 = a[5] < b[5] ? a[5] : b[5];
  float d0 = hi0 - lo0, d1 = hi1 - lo1, d2 = hi2 - lo2;
  d0 = d0 < 0.f ? 0.f : d0;
  d1 = d1 < 0.f ? 0.f : d1;
  d2 = d2 < 0.f ? 0.f : d2;
  float inter = (d0 * d1) * d2;
  /* utils.py:148-149 */
  float uni = (va + vb) - inter;
  return inter / uni;
}

static inline int level_of(float vol) {
  if (!(vol > 0.f)) return 0;
  int e;
  frexpf(vol, &e); /* vol = m * 2^e, m in [0.5, 1) */
  e += 64;
  if (e < 1) e = 1;
  if (e > MAX_LEVELS - 1) e = MAX_LEVELS - 1;
  return e;
}

static inline int cell_of(const level_t* L, int axis, float x) {
  int c = (int)floorf((x - L->lo[axis]) / L->cell);
  if (c < 0) c = 0;
  if (c > L->dim[axis] - 1) c = L->dim[axis] - 1;
  return c;
}

/* boxes: n x 6 fp32 [x0 y0 z0 x1 y1 z1], score-sorted.  keep: n bytes.  Returns the kept count, -1 on
 * allocation failure, -2 on a precondition violation. */
long long nms_oracle_greedy(const float* boxes, long long n, float thr, uint8_t* keep) {
  if (n <= 0) return 0;
  if (!(thr >= 0.f)) return -2;
  float* vol = (float*)malloc((size_t)n * sizeof(float));
  int32_t* next = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  uint8_t* lvl = (uint8_t*)malloc((size_t)n);
  level_t* L = (level_t*)calloc(MAX_LEVELS, sizeof(level_t));
  if (!vol || !next || !lvl || !L) return -1;
  float glo[3] = {INFINITY, INFINITY, INFINITY}, ghi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (long long i = 0; i < n; ++i) {
    const float* b = boxes + 6 * i;
    for (int k = 0; k < 6; ++k)
      if (!isfinite(b[k])) return -2;
    if (b[3] < b[0] || b[4] < b[1] || b[5] < b[2]) return -2;
    vol[i] = ((b[3] - b[0]) * (b[4] - b[1])) * (b[5] - b[2]); /* utils.py:142-147 */
    const int lv = level_of(vol[i]);
    lvl[i] = (uint8_t)lv;
    L[lv].used = 1;
    for (int k = 0; k < 3; ++k) {
      const float e = b[3 + k] - b[k];
      if (e > L[lv].maxext) L[lv].maxext = e;
      if (b[k] < glo[k]) glo[k] = b[k];
      if (b[k] > ghi[k]) ghi[k] = b[k];
    }
  }
  for (int lv = 0; lv < MAX_LEVELS; ++lv) {
    if (!L[lv].used) continue;
    float cell = L[lv].maxext > 0.f ? 0.5f * L[lv].maxext : 1e-6f; /* any cell size is correct; this one is fast */
    for (int k = 0; k < 3; ++k) {
      const float need = (ghi[k] - glo[k]) / (float)(MAX_DIM - 2);
      if (need > cell) cell = need;
    }
    L[lv].cell = cell;
    size_t cells = 1;
    for (int k = 0; k < 3; ++k) {
      L[lv].lo[k] = glo[k];
      L[lv].dim[k] = (int)floorf((ghi[k] - glo[k]) / cell) + 2;
      cells *= (size_t)L[lv].dim[k];
    }
    L[lv].head = (int32_t*)malloc(cells * sizeof(int32_t));
    if (!L[lv].head) return -1;
    memset(L[lv].head, 0xff, cells * sizeof(int32_t));
  }
  const int use_vol = thr > 0.f;
  long long kept = 0;
  for (long long i = 0; i < n; ++i) {
    const float* q = boxes + 6 * i;
    const float vq = vol[i];
    int suppressed = 0;
    /* volume window of possible suppressors, widened by a relative 1e-4 so that rounding can never drop a pair */
    const float vmin = use_vol ? vq * thr * (1.f - 1e-4f) : -INFINITY;
    const float vmax = use_vol ? (vq / thr) * (1.f + 1e-4f) : INFINITY;
    for (int lv = 0; lv < MAX_LEVELS && !suppressed; ++lv) {
      const level_t* G = &L[lv];
      if (!G->used) continue;
      if (use_vol && lv > 1 && lv < MAX_LEVELS - 1) {
        /* level lv holds volumes in [2^(lv-65), 2^(lv-64)); the two end levels also collect what is beyond */
        const float lv_lo = ldexpf(1.f, lv - 65), lv_hi = ldexpf(1.f, lv - 64);
        if (lv_hi < vmin || lv_lo > vmax) continue;
      }
      int c0[3], c1[3];
      for (int k = 0; k < 3; ++k) {
        c0[k] = cell_of(G, k, q[k] - G->maxext);
        c1[k] = cell_of(G, k, q[3 + k]);
      }
      for (int a = c0[0]; a <= c1[0] && !suppressed; ++a)
        for (int b = c0[1]; b <= c1[1] && !suppressed; ++b)
          for (int c = c0[2]; c <= c1[2] && !suppressed; ++c) {
            int32_t j = G->head[((size_t)a * G->dim[1] + b) * G->dim[2] + c];
            while (j >= 0) {
              const float vj = vol[j];
              if (vj >= vmin && vj <= vmax) {
                if (iou_ref(boxes + 6 * (long long)j, vj, q, vq) > thr) { /* ssd3d.py:422, strict */
                  suppressed = 1;
                  break;
                }
              }
              j = next[j];
            }
          }
    }
    keep[i] = (uint8_t)!suppressed;
    if (!suppressed) {
      ++kept;
      level_t* G = &L[lvl[i]];
      const size_t cidx = ((size_t)cell_of(G, 0, q[0]) * G->dim[1] + cell_of(G, 1, q[1])) * G->dim[2] + cell_of(G, 2, q[2]);
      next[i] = G->head[cidx];
      G->head[cidx] = (int32_t)i;
    }
  }
  for (int lv = 0; lv < MAX_LEVELS; ++lv) free(L[lv].head);
  free(L);
  free(lvl);
  free(next);
  free(vol);
  return kept;
}
