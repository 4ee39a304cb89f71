// Result hand-off: fixed-shape detections -> COCO detection records -> the JSON file the
// reference writes.
//
// Reference: maskrcnn_benchmark/data/datasets/evaluation/coco/coco_eval.py:70-176 (prepare_for_coco_detection): per
// image  prediction.resize((image_width, image_height))  (structures/bounding_box.py:91-127),  .convert("xywh")
// (:55-73, TO_REMOVE = 1),  bbox.tolist() / scores.tolist(),  one dict per box with keys image_id, category_id, bbox,
// score (:146-156),  json.dump(coco_results, f, sort_keys=True, indent=4, separators=(',', ':'))  (:163-165).
//
// Device half (osd_coco_records): one kernel turns the padded [E,K,4] / [E,K] / [E] output of the post-processing
// stage into compact (x, y, w, h, score) rows in episode order -- the resize and xywh arithmetic are single rounded
// fp32 operations in the reference's order, so the values are bit-identical.
// Host half (osd_coco_write_json): formats the rows exactly like CPython's json.dump with the reference's arguments
// (float repr = shortest round-trip digits, fixed notation for 1e-4 <= |x| < 1e16), so the file is byte-identical.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "osd_common.cuh"
#include "osd_device_utils.cuh"

namespace osd {
namespace {

struct CocoArgs {
  const float4* boxes;     // [E, K]
  const float* scores;     // [E, K]
  const int32_t* count;    // [E]
  const int32_t* det_wh;   // [E, 2] BoxList.size of the detections (w, h)
  const int32_t* orig_wh;  // [E, 2] img_info width, height
  int E, K;
  float* rec;              // [sum count, 5]
  int32_t* rec_episode;    // [sum count]
  int32_t* total;          // [1]
};

// One CTA per episode; its first row is the sum of the counts before it (E is small).
__global__ void __launch_bounds__(256) coco_records_kernel(CocoArgs A) {
  __shared__ int warp_tot[33];
  const int e = blockIdx.x, tid = threadIdx.x;
  int part = 0;
  for (int i = tid; i < e; i += blockDim.x) part += min(max(A.count[i], 0), A.K);
  const int base = block_sum(part, warp_tot);
  const int n = min(max(A.count[e], 0), A.K);
  if (e == A.E - 1 && tid == 0) *A.total = base + n;
  // bounding_box.py:99: ratios = float(s) / float(s_orig) in double; the tensor product rounds the scalar to fp32
  const double rw_d = (double)A.orig_wh[2 * e] / (double)A.det_wh[2 * e];
  const double rh_d = (double)A.orig_wh[2 * e + 1] / (double)A.det_wh[2 * e + 1];
  const float rw = (float)rw_d, rh = (rw_d == rh_d) ? (float)rw_d : (float)rh_d;   // :100-103 equal ratios: one factor
  for (int i = tid; i < n; i += blockDim.x) {
    const float4 b = A.boxes[(size_t)e * A.K + i];
    const float x1 = __fmul_rn(b.x, rw), y1 = __fmul_rn(b.y, rh), x2 = __fmul_rn(b.z, rw), y2 = __fmul_rn(b.w, rh);
    float* r = A.rec + (size_t)(base + i) * 5;
    r[0] = x1;
    r[1] = y1;
    r[2] = __fadd_rn(__fsub_rn(x2, x1), 1.0f);   // bounding_box.py:67-70
    r[3] = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
    r[4] = A.scores[(size_t)e * A.K + i];
    A.rec_episode[base + i] = e;
  }
}

// CPython float.__repr__ (format code 'r'): shortest digits that round-trip; exponent form iff decpt <= -4 or decpt > 16
void py_float_repr(double v, std::string& out) {
  if (std::isnan(v)) { out += "NaN"; return; }                 // json.dump(allow_nan=True)
  if (std::isinf(v)) { out += v < 0 ? "-Infinity" : "Infinity"; return; }
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
  std::string s(buf, res.ptr);                                  // [-]d[.ddd]e[+-]XX
  size_t pos = 0;
  if (s[0] == '-') { out += '-'; pos = 1; }
  const size_t epos = s.find('e');
  std::string digits;
  for (size_t i = pos; i < epos; ++i)
    if (s[i] != '.') digits += s[i];
  const int exp10 = atoi(s.c_str() + epos + 1);
  const int decpt = exp10 + 1;
  const int nd = (int)digits.size();
  if (digits == "0") { out += "0.0"; return; }
  if (decpt <= -4 || decpt > 16) {
    out += digits[0];
    if (nd > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
    char eb[16];
    snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
    out += eb;
  } else if (decpt <= 0) {
    out += "0.";
    out.append((size_t)(-decpt), '0');
    out += digits;
  } else if (decpt >= nd) {
    out += digits;
    out.append((size_t)(decpt - nd), '0');
    out += ".0";
  } else {
    out.append(digits, 0, (size_t)decpt);
    out += '.';
    out.append(digits, (size_t)decpt, std::string::npos);
  }
}

}  // namespace
}  // namespace osd

extern "C" int osd_coco_records(const float* boxes, const float* scores, const int32_t* count, const int32_t* det_wh,
                                const int32_t* orig_wh, int32_t num_episodes, int32_t rows_per_episode, float* records,
                                int32_t* record_episode, int32_t* total, void* stream_) {
  using namespace osd;
  OSD_REQUIRE(num_episodes >= 0 && num_episodes <= 65535 && rows_per_episode >= 0, "osd_coco_records: bad sizes");
  OSD_REQUIRE(total != nullptr, "osd_coco_records: total is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (num_episodes == 0) {
    OSD_CUDA(cudaMemsetAsync(total, 0, sizeof(int32_t), stream));
    return OSD_OK;
  }
  OSD_REQUIRE(boxes && scores && count && det_wh && orig_wh && records && record_episode, "osd_coco_records: null pointer");
  OSD_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "osd_coco_records: boxes must be 16-byte aligned");
  CocoArgs A{};
  A.boxes = reinterpret_cast<const float4*>(boxes);
  A.scores = scores;
  A.count = count;
  A.det_wh = det_wh;
  A.orig_wh = orig_wh;
  A.E = num_episodes;
  A.K = rows_per_episode;
  A.rec = records;
  A.rec_episode = record_episode;
  A.total = total;
  coco_records_kernel<<<num_episodes, 256, 0, stream>>>(A);
  OSD_LAUNCH_CHECK("coco_records_kernel");
  return OSD_OK;
}

extern "C" int osd_coco_write_json(const float* records, const int32_t* record_episode, int64_t num_records,
                                   const int64_t* image_ids, const int64_t* category_ids, int32_t num_episodes,
                                   const char* path) {
  using namespace osd;
  OSD_REQUIRE(path != nullptr, "osd_coco_write_json: path is null");
  OSD_REQUIRE(num_records >= 0, "osd_coco_write_json: negative record count");
  OSD_REQUIRE(num_records == 0 || (records && record_episode && image_ids && category_ids), "osd_coco_write_json: null pointer");
  std::string out;
  out.reserve((size_t)num_records * 220 + 16);
  if (num_records == 0) {
    out = "[]";   // json.dump([], indent=4) writes "[]"
  } else {
    out += "[\n";
    char ib[32];
    for (int64_t i = 0; i < num_records; ++i) {
      const float* r = records + i * 5;
      const int32_t e = record_episode[i];
      OSD_REQUIRE(e >= 0 && e < num_episodes, "osd_coco_write_json: record %lld names episode %d of %d", (long long)i, e,
                  num_episodes);
      out += "    {\n        \"bbox\":[\n";
      for (int k = 0; k < 4; ++k) {
        out += "            ";
        py_float_repr((double)r[k], out);
        out += k < 3 ? ",\n" : "\n";
      }
      out += "        ],\n        \"category_id\":";
      snprintf(ib, sizeof(ib), "%lld", (long long)category_ids[e]);
      out += ib;
      out += ",\n        \"image_id\":";
      snprintf(ib, sizeof(ib), "%lld", (long long)image_ids[e]);
      out += ib;
      out += ",\n        \"score\":";
      py_float_repr((double)r[4], out);
      out += i + 1 < num_records ? "\n    },\n" : "\n    }\n";
    }
    out += "]";
  }
  FILE* f = fopen(path, "wb");
  if (!f) {
    set_error("osd_coco_write_json: cannot open '%s' for writing", path);
    return OSD_ERR_INVALID;
  }
  const size_t w = fwrite(out.data(), 1, out.size(), f);
  const int rc = fclose(f);
  if (w != out.size() || rc != 0) {
    set_error("osd_coco_write_json: short write to '%s'", path);
    return OSD_ERR_INVALID;
  }
  return OSD_OK;
}
