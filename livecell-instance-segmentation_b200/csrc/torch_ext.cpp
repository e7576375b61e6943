// torch_ext.cpp — the thin torch extension above the C-ABI (include/lcr.h): C++ entry points for the three
// torchvision operators the reference imports (src/custom_maskrcnn.py:5) on the small, launch-latency-bound
// shapes of the training step (BASELINE config C2: K = 128 RoIs on a 64x64 map) and of single-frame inference.
//
// Nothing is computed here: every function checks its tensors, allocates outputs through the caching allocator and
// calls liblcr.so (hand-written sm_100a kernels) on the current CUDA stream.  What it removes is the Python call
// path (ctypes marshalling + torch.autograd.Function.apply + per-call tensor plumbing, ~0.25 ms per RoIAlign
// forward+backward, twice torchvision's C++ dispatcher) — the autograd node of RoIAlign lives in C++.
//
//   roi_align(input, rois | [boxes], spatial_scale, PH, PW, sampling_ratio, aligned) -> [K, C, PH, PW]
//       torchvision::roi_align + torchvision::_roi_align_backward (TV:ops/roi_align.py:204-260), differentiable
//       w.r.t. `input`; list-of-boxes form of TV:ops/_utils.py:18-25.
//   nms(boxes, scores, iou_threshold) -> int64 [M]          torchvision::nms (TV:ops/boxes.py:20-48)
//
// CPU tensors are rejected: there is no CPU fallback.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <vector>

#include "lcr.h"

namespace {

void check_rc(int rc, const char* what) {
  if (rc != LCR_OK) {
    std::string msg = std::string("liblcr ") + what + ": " + lcr_error_string(rc);
    if (rc == LCR_ERR_CUDA) msg += " (cudaError " + std::to_string(lcr_last_cuda_error()) + ")";
    TORCH_CHECK(false, msg);
  }
}

void* current_stream() { return reinterpret_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

bool nhwc_dense(const at::Tensor& t) {
  const int64_t C = t.size(1), H = t.size(2), W = t.size(3);
  return t.stride(0) == H * W * C && t.stride(1) == 1 && t.stride(2) == W * C && t.stride(3) == C;
}

LcrFeatLevel level_of(const at::Tensor& t, double scale) {
  LcrFeatLevel lv{};
  lv.data = t.data_ptr<float>();
  lv.N = (int)t.size(0);
  lv.H = (int)t.size(2);
  lv.W = (int)t.size(3);
  lv.sn = t.stride(0);
  lv.sc = t.stride(1);
  lv.sh = t.stride(2);
  lv.sw = t.stride(3);
  lv.spatial_scale = (float)scale;
  return lv;
}

at::Tensor as_rois(const at::Tensor& rois) {
  TORCH_CHECK(rois.is_cuda(), "liblcr ops need CUDA tensors: the region pipeline has no CPU fallback");
  TORCH_CHECK(rois.dim() == 2 && rois.size(1) == 5, "rois must be [K, 5]");
  return rois.scalar_type() == at::kFloat ? rois.contiguous() : rois.to(at::kFloat).contiguous();
}

// list[Tensor[K_i, 4]] -> Tensor[K, 5] with the image index in column 0 (TV:ops/_utils.py:18-25)
at::Tensor boxes_to_rois(const std::vector<at::Tensor>& boxes) {
  TORCH_CHECK(!boxes.empty(), "roi_align: empty box list");
  if (boxes.size() == 1) return at::constant_pad_nd(boxes[0], {1, 0}, 0.0);  // the reference's call form: one image
  std::vector<at::Tensor> parts;
  parts.reserve(boxes.size());
  for (size_t i = 0; i < boxes.size(); ++i) parts.push_back(at::constant_pad_nd(boxes[i], {1, 0}, (double)i));
  return at::cat(parts, 0);
}

struct RoIAlignFn : public torch::autograd::Function<RoIAlignFn> {
  static at::Tensor forward(torch::autograd::AutogradContext* ctx, const at::Tensor& input, const at::Tensor& rois_in,
                            double scale, int64_t PH, int64_t PW, int64_t sr, int64_t flags) {
    TORCH_CHECK(input.is_cuda(), "liblcr ops need CUDA tensors: the region pipeline has no CPU fallback");
    TORCH_CHECK(input.dim() == 4, "roi_align: input must be [N, C, H, W]");
    const c10::cuda::CUDAGuard guard(input.device());
    const at::Tensor rois = as_rois(rois_in);
    const int64_t N = input.size(0), C = input.size(1), H = input.size(2), W = input.size(3), K = rois.size(0);
    at::Tensor feat = input.scalar_type() == at::kFloat ? input : input.to(at::kFloat);
    if (!nhwc_dense(feat)) {
      feat = feat.contiguous();
      // NCHW input: transposing the map once pays off above ~100 RoIs per 130x176 map (the warp-item kernels gather whole
      // channel vectors; below that roi_fwd_planes_kernel pools the planes directly) — same rule as roi_align._prefer_nhwc,
      // measured in profiles/r02c_roi_nchw_exp.jsonl
      if (sr == 2 && PH == PW && (PH == 7 || PH == 14) && C % 4 == 0 &&
          (double)K * PH * PW * C >= 8e5 + 0.08 * (double)N * C * H * W) {
        at::Tensor t = at::empty_strided({N, C, H, W}, {H * W * C, 1, W * C, C}, feat.options());
        check_rc(lcr_nchw_to_nhwc_f32(feat.data_ptr<float>(), t.data_ptr<float>(), (int)N, (int)C, (int)H, (int)W, current_stream()),
                 "nchw_to_nhwc");
        feat = t;
      }
    }
    at::Tensor out = at::empty({K, C, PH, PW}, feat.options());
    if (K > 0) {
      const LcrFeatLevel lv = level_of(feat, scale);
      check_rc(lcr_roi_align_fwd_f32(&lv, 1, (int)C, rois.data_ptr<float>(), nullptr, (int)K, (int)PH, (int)PW, (int)sr,
                                     (int)flags, out.data_ptr<float>(), current_stream()),
               "roi_align_fwd");
    }
    ctx->save_for_backward({rois});
    ctx->saved_data["shape"] = std::vector<int64_t>{N, C, H, W};
    ctx->saved_data["scale"] = scale;
    ctx->saved_data["sr"] = sr;
    ctx->saved_data["flags"] = flags;   // bit 0 aligned, bit 1 CPU-op coordinate rounding (include/lcr.h)
    return out;
  }

  static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grads) {
    const at::Tensor rois = ctx->get_saved_variables()[0];
    const std::vector<int64_t> shape = ctx->saved_data["shape"].toIntVector();
    const int64_t N = shape[0], C = shape[1], H = shape[2], W = shape[3];
    const c10::cuda::CUDAGuard guard(rois.device());
    at::Tensor g = grads[0];
    g = (g.scalar_type() == at::kFloat ? g : g.to(at::kFloat)).contiguous();
    const int64_t K = g.size(0), PH = g.size(2), PW = g.size(3);
    // channels_last gradient map: the backward kernels scatter whole channel vectors
    at::Tensor gin = at::empty_strided({N, C, H, W}, {H * W * C, 1, W * C, C}, g.options());
    const LcrFeatLevel lv = level_of(gin, ctx->saved_data["scale"].toDouble());
    check_rc(lcr_roi_align_bwd_f32(K > 0 ? g.data_ptr<float>() : nullptr, &lv, 1, (int)C, rois.data_ptr<float>(), nullptr, (int)K,
                                   (int)PH, (int)PW, (int)ctx->saved_data["sr"].toInt(), (int)ctx->saved_data["flags"].toInt(),
                                   /*zero_grad=*/1, current_stream()),
             "roi_align_bwd");
    return {gin, at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor(), at::Tensor()};
  }
};

at::Tensor roi_align(const at::Tensor& input, const at::Tensor& rois, double scale, int64_t PH, int64_t PW, int64_t sr, int64_t flags) {
  return RoIAlignFn::apply(input, rois, scale, PH, PW, sr < 0 ? 0 : sr, flags);
}

at::Tensor roi_align_list(const at::Tensor& input, const std::vector<at::Tensor>& boxes, double scale, int64_t PH, int64_t PW,
                          int64_t sr, int64_t flags) {
  return RoIAlignFn::apply(input, boxes_to_rois(boxes), scale, PH, PW, sr < 0 ? 0 : sr, flags);
}

// torchvision.ops.nms: int64 indices of the kept boxes in decreasing score order.  One host sync (the dense return
// type needs the count), as in torchvision's own CUDA op.
at::Tensor nms(const at::Tensor& boxes_in, const at::Tensor& scores_in, double iou_threshold) {
  TORCH_CHECK(boxes_in.is_cuda() && scores_in.is_cuda(), "liblcr ops need CUDA tensors: the region pipeline has no CPU fallback");
  TORCH_CHECK(boxes_in.dim() == 2 && boxes_in.size(1) == 4 && scores_in.dim() == 1 && scores_in.size(0) == boxes_in.size(0),
              "nms: boxes [N, 4], scores [N]");
  const c10::cuda::CUDAGuard guard(boxes_in.device());
  const int64_t n = boxes_in.size(0);
  auto opt_i64 = boxes_in.options().dtype(at::kLong);
  if (n == 0) return at::empty({0}, opt_i64);
  const at::Tensor boxes = (boxes_in.scalar_type() == at::kFloat ? boxes_in : boxes_in.to(at::kFloat)).contiguous();
  const at::Tensor scores = (scores_in.scalar_type() == at::kFloat ? scores_in : scores_in.to(at::kFloat)).contiguous();
  const size_t ws_bytes = lcr_nms_workspace_bytes(1, (int)n);
  at::Tensor ws = at::empty({(int64_t)ws_bytes}, boxes.options().dtype(at::kByte));
  at::Tensor keep = at::empty({n}, opt_i64);
  at::Tensor count = at::empty({1}, boxes.options().dtype(at::kInt));
  check_rc(lcr_nms_f32(boxes.data_ptr<float>(), scores.data_ptr<float>(), nullptr, nullptr, 1, (int)n, iou_threshold, 0.f, 0, (int)n,
                       keep.data_ptr<int64_t>(), count.data_ptr<int>(), ws.data_ptr(), ws_bytes, current_stream()),
           "nms");
  const int64_t m = count.item<int>();
  return keep.narrow(0, 0, m);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "liblcr torch extension: C++ call path for roi_align (autograd) and nms over the C-ABI of include/lcr.h";
  m.def("roi_align", &roi_align, "RoIAlign forward (differentiable w.r.t. input), rois [K,5]");
  m.def("roi_align_list", &roi_align_list, "RoIAlign forward, list of per-image boxes [K_i,4]");
  m.def("nms", &nms, "greedy NMS keep list (int64, score order)");
  m.def("lcr_version", []() { return lcr_version(); });
}
