/* CPU oracle for greedy NMS -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C restatement of the reference's two NMS back-ends:
 *   cmp_mode 1 ('>=')  mmdet/ops/nms/src/nms_cpu.cpp:4-59   (suppress if ovr >= thr, :55)
 *   cmp_mode 0 ('>')   mmdet/ops/nms/src/nms_kernel.cu:13-67,105-123 (suppress if IoU > thr, :60)
 * Both use the "+1 pixel" box convention and the same fp32 operation order
 * (nms_cpu.cpp:18,47-54; nms_kernel.cu:13-21), so one function serves both.
 *
 * The reference sorts with at::sort (nms_cpu.cpp:20, nms_kernel.cu:74), whose
 * order among equal scores is unspecified; this oracle uses a stable sort
 * (ties keep ascending original index) and the parity tests use distinct
 * scores.  Output: ascending original indices of the kept boxes
 * (nms_cpu.cpp:58 nonzero; nms_kernel.cu:127-130 final sort).
 *
 * Pinned against the reference's own nms_cpu.cpp compiled unmodified into
 * oracle/_ref (tests/test_oracle_cpu.py::test_nms_oracle_matches_reference_cpu).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC nms_oracle.c -o _build/libnms_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int64_t idx; } kv_t;

static void merge_sort_desc(kv_t *a, kv_t *tmp, int64_t n) {
  if (n < 2) return;
  int64_t h = n / 2;
  merge_sort_desc(a, tmp, h);
  merge_sort_desc(a + h, tmp, n - h);
  int64_t i = 0, j = h, k = 0;
  while (i < h && j < n) {
    /* stable: take from the left run unless the right is strictly greater */
    if (a[j].score > a[i].score) tmp[k++] = a[j++]; else tmp[k++] = a[i++];
  }
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(kv_t));
}

/* dets: [n,5] (x1,y1,x2,y2,score) fp32.  keep: [n] out.  Returns number kept. */
int64_t kgdet_oracle_nms(const float *dets, int64_t n, float thr, int cmp_mode,
                         int64_t *keep) {
  if (n <= 0) return 0;
  kv_t *order = (kv_t *)malloc((size_t)n * sizeof(kv_t));
  kv_t *tmp = (kv_t *)malloc((size_t)n * sizeof(kv_t));
  float *area = (float *)malloc((size_t)n * sizeof(float));
  uint8_t *sup = (uint8_t *)calloc((size_t)n, 1);
  for (int64_t i = 0; i < n; ++i) {
    const float *b = dets + 5 * i;
    order[i].score = b[4];
    order[i].idx = i;
    area[i] = (b[2] - b[0] + 1) * (b[3] - b[1] + 1); /* nms_cpu.cpp:18 */
  }
  merge_sort_desc(order, tmp, n);
  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = order[_i].idx;
    if (sup[i]) continue;
    const float *bi = dets + 5 * i;
    float iarea = area[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j].idx;
      if (sup[j]) continue;
      const float *bj = dets + 5 * j;
      float xx1 = bi[0] > bj[0] ? bi[0] : bj[0];
      float yy1 = bi[1] > bj[1] ? bi[1] : bj[1];
      float xx2 = bi[2] < bj[2] ? bi[2] : bj[2];
      float yy2 = bi[3] < bj[3] ? bi[3] : bj[3];
      float w = xx2 - xx1 + 1; if (!(w > 0.f)) w = 0.f;
      float h = yy2 - yy1 + 1; if (!(h > 0.f)) h = 0.f;
      float inter = w * h;
      float ovr = inter / (iarea + area[j] - inter);
      if (cmp_mode ? (ovr >= thr) : (ovr > thr)) sup[j] = 1;
    }
  }
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) if (!sup[i]) keep[m++] = i;
  free(order); free(tmp); free(area); free(sup);
  return m;
}
