// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// pybind glue that exposes the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/pet/lib/ops/csrc, never copied into this repo) as a python module so the
// oracle restatement in oracle/ can be pinned against the reference's own arithmetic.
//   roi_align_forward / roi_align_backward -> ROIAlign/ROIAlign.h:57-146 (CPU + CUDA dispatch)
//   soft_nms                               -> NMS/soft_nms.h:19-40      (CPU, hard mode = second NMS oracle)
//   ml_nms                                 -> NMS/ml_nms.h:16-39        (CUDA only; WITH_CUDA build)
#include <torch/extension.h>
#include "ROIAlign/ROIAlign.h"
#include "NMS/soft_nms.h"
#ifdef WITH_CUDA
#include "NMS/ml_nms.h"
#endif

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("roi_align_forward", &pet::ROIAlign_forward);
  m.def("roi_align_backward", &pet::ROIAlign_backward);
  m.def("soft_nms", &pet::soft_nms);
#ifdef WITH_CUDA
  m.def("ml_nms", &pet::ml_nms);
#endif
}
