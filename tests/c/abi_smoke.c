/* The C ABI from plain C (no C++, no Python, no torch): the header must compile as C99, the library must link, and the
 * entry points that need no GPU (version, plans, argument validation, the JSON writer) must work.
 * Built and run by tests/test_abi_cpu.py::test_header_is_plain_c_and_links_from_c. */
#include <stdio.h>
#include <string.h>

#include "osd_b200.h"

int main(int argc, char** argv) {
  if (osd_version() <= 0) return 1;
  osd_nms_plan np;
  if (osd_batched_nms_plan(16, 11600, &np) != OSD_OK || np.workspace_bytes == 0 || np.padded_len % 64 != 0) return 2;
  osd_fcos_config fc;
  memset(&fc, 0, sizeof fc);
  fc.num_levels = 2; fc.batch = 4;
  fc.height[0] = 100; fc.width[0] = 168; fc.stride[0] = 8;
  fc.height[1] = 50;  fc.width[1] = 84;  fc.stride[1] = 16;
  fc.pre_nms_thresh = 0.0f; fc.pre_nms_top_n = 6000; fc.nms_thresh = 0.8f; fc.post_nms_top_n = 2000; fc.early_exit = 1;
  osd_fcos_plan fp;
  if (osd_fcos_postprocess_plan(&fc, &fp) != OSD_OK || fp.cand_capacity != 6000 + 4200 || fp.out_capacity != 2000) return 3;
  fc.num_levels = 99;
  if (osd_fcos_postprocess_plan(&fc, &fp) == OSD_OK || strlen(osd_last_error()) == 0) return 4;
  osd_box_post_config bc;
  memset(&bc, 0, sizeof bc);
  bc.batch = 2; bc.rois_per_image = 100; bc.num_logits = 2; bc.reg_columns = 8; bc.reg_offset = 4;
  bc.score_mode = OSD_SCORE_SOFTMAX;
  bc.weights[0] = bc.weights[1] = 10.f; bc.weights[2] = bc.weights[3] = 5.f;
  osd_box_post_plan bp;
  if (osd_box_postprocess_plan(&bc, &bp) != OSD_OK || bp.out_capacity != 100) return 5;
  if (argc > 1) {
    const float rec[5] = {1.5f, 2.0f, 3.25f, 4.0f, 0.1f};
    const int32_t ep[1] = {0};
    const int64_t img[1] = {7}, cat[1] = {3};
    if (osd_coco_write_json(rec, ep, 1, img, cat, 1, argv[1]) != OSD_OK) return 6;
  }
  printf("c-abi ok, version %d\n", osd_version());
  return 0;
}
