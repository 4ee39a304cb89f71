// TEST INFRASTRUCTURE ONLY -- never imported by the product path.
//
// Builds the *unmodified* reference CPU operators for the hot path straight from
// /root/reference (the sources are #included where they lie; nothing is copied
// into this repository).  Output goes to oracle/_ref/osd_ref_C*.so (git-ignored,
// shipped to the GPU box by gpurun).
//
// The reference targets torch 1.0/1.1: its AT_DISPATCH_FLOATING_TYPES call sites
// pass `tensor.type()` (a DeprecatedTypeProperties) where torch 2.11's dispatch
// macro expects a ScalarType (maskrcnn_benchmark/csrc/cpu/nms_cpu.cpp:71,
// maskrcnn_benchmark/csrc/cpu/ROIAlign_cpu.cpp:242).  The macro resolves the type
// through `::detail::scalar_type(the_type)`, so one extra overload declared
// *before* the reference sources are included makes them compile untouched.
#include <torch/extension.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) {
  return t.scalarType();
}
}  // namespace detail

// Reference sources by absolute, quoted path: build_ref.py passes
//   -DOSD_REF_NMS_CPU_CPP='"<ref>/maskrcnn_benchmark/csrc/cpu/nms_cpu.cpp"' etc.
// and -I<ref>/maskrcnn_benchmark/csrc so their own `#include "cpu/vision.h"` resolves.
#include OSD_REF_NMS_CPU_CPP       // nms_cpu_kernel / nms_cpu      (csrc/cpu/nms_cpu.cpp:5-75)
#include OSD_REF_ROIALIGN_CPU_CPP  // ROIAlign_forward_cpu          (csrc/cpu/ROIAlign_cpu.cpp:113-257)
#include OSD_REF_NMS_H             // nms() device dispatcher       (csrc/nms.h:10-28)
#include OSD_REF_ROIALIGN_H        // ROIAlign_forward dispatcher   (csrc/ROIAlign.h:11-25)

// Same export names as the reference's pybind module (csrc/vision.cpp:7-15), restricted
// to the operators that exist on CPU.
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("nms", &nms, "non-maximum suppression (reference nms_cpu)");
  m.def("roi_align_forward", &ROIAlign_forward, "ROIAlign_forward (reference CPU)");
}
