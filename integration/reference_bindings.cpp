// INTEGRATION.md section B as compilable code: the bodies a maintainer of cv-stuttgart/PCFA would put into the
// reference's pybind11 modules so that they call libpcfa_b200.so (include/pcfa_b200.h) instead of their own kernels.
//   spatial_correlation_sampler_backend : models/PWCNet/cpu_spatial_correlation_sampler-0.3.0/Correlation_Module/correlation_sampler.cpp:58-117
//   correlation_cuda                    : models/FlowNet/correlation_package/correlation_cuda.cc:10-171
//   resample2d_cuda                     : models/FlowNet/resample2d_package/resample2d_cuda.cc:6-31
//   channelnorm_cuda                    : models/FlowNet/channelnorm_package/channelnorm_cuda.cc:6-30
// Same function names, argument lists and return conventions as those files.  tests/test_abi.py compiles this file
// with -fsyntax-only against the torch headers, so the prose in INTEGRATION.md cannot drift from the C ABI.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <vector>
#include "pcfa_b200.h"

#define PCFA_CHECK_INPUT(x) TORCH_CHECK((x).is_cuda() && (x).is_contiguous() && (x).scalar_type() == at::kFloat, #x " must be a contiguous CUDA float tensor")
static inline pcfa_stream_t cur_stream() { return (pcfa_stream_t)at::cuda::getCurrentCUDAStream().stream(); }
static inline void ok(int st, const char* what) { TORCH_CHECK(st == PCFA_OK, what, ": ", pcfa_status_string(st)); }

// ------------------------------------------------------------------ spatial_correlation_sampler_backend
torch::Tensor correlation_sample_forward(torch::Tensor input1, torch::Tensor input2, int kH, int kW, int patchH, int patchW,
                                         int padH, int padW, int dilationH, int dilationW, int dilation_patchH,
                                         int dilation_patchW, int dH, int dW) {
  PCFA_CHECK_INPUT(input1); PCFA_CHECK_INPUT(input2);
  pcfa_scs_params p{kH, kW, patchH, patchW, padH, padW, dilationH, dilationW, dilation_patchH, dilation_patchW, dH, dW};
  int oH = 0, oW = 0;
  ok(pcfa_scs_output_size((int)input1.size(2), (int)input1.size(3), &p, &oH, &oW), "pcfa_scs_output_size");
  auto out = torch::empty({input1.size(0), patchH, patchW, oH, oW}, input1.options());
  ok(pcfa_scs_forward(input1.data_ptr<float>(), input2.data_ptr<float>(), out.data_ptr<float>(), (int)input1.size(0),
                      (int)input1.size(1), (int)input1.size(2), (int)input1.size(3), &p, 1.0f, cur_stream()), "pcfa_scs_forward");
  return out;
}

std::vector<torch::Tensor> correlation_sample_backward(torch::Tensor input1, torch::Tensor input2, torch::Tensor grad_output,
                                                       int kH, int kW, int patchH, int patchW, int padH, int padW,
                                                       int dilationH, int dilationW, int dilation_patchH, int dilation_patchW,
                                                       int dH, int dW) {
  PCFA_CHECK_INPUT(input1); PCFA_CHECK_INPUT(input2);
  grad_output = grad_output.contiguous();
  pcfa_scs_params p{kH, kW, patchH, patchW, padH, padW, dilationH, dilationW, dilation_patchH, dilation_patchW, dH, dW};
  auto g1 = torch::empty_like(input1), g2 = torch::empty_like(input2);
  ok(pcfa_scs_backward(input1.data_ptr<float>(), input2.data_ptr<float>(), grad_output.data_ptr<float>(), g1.data_ptr<float>(),
                       g2.data_ptr<float>(), (int)input1.size(0), (int)input1.size(1), (int)input1.size(2), (int)input1.size(3), &p,
                       1.0f, cur_stream()), "pcfa_scs_backward");
  return {g1, g2};
}

// ------------------------------------------------------------------ correlation_cuda (FlowNet2)
int correlation_forward_cuda(at::Tensor& input1, at::Tensor& input2, at::Tensor& rInput1, at::Tensor& rInput2, at::Tensor& output,
                             int pad_size, int kernel_size, int max_displacement, int stride1, int stride2, int /*corr_type_multiply*/) {
  int oc = 0, oh = 0, ow = 0;
  ok(pcfa_fn2corr_output_size((int)input1.size(2), (int)input1.size(3), pad_size, kernel_size, max_displacement, stride1, stride2,
                              &oc, &oh, &ow), "pcfa_fn2corr_output_size");
  output.resize_({input1.size(0), oc, oh, ow});            // rInput1 / rInput2 stay empty: no padded NHWC copies are built
  int st = pcfa_fn2corr_forward(input1.data_ptr<float>(), input2.data_ptr<float>(), output.data_ptr<float>(), (int)input1.size(0),
                                (int)input1.size(1), (int)input1.size(2), (int)input1.size(3), pad_size, kernel_size,
                                max_displacement, stride1, stride2, cur_stream());
  if (st != PCFA_OK) AT_ERROR("correlation_forward_cuda: ", pcfa_status_string(st));     // correlation_cuda.cc:80-83
  return 1;
}

int correlation_backward_cuda(at::Tensor& input1, at::Tensor& input2, at::Tensor& rInput1, at::Tensor& rInput2, at::Tensor& gradOutput,
                              at::Tensor& gradInput1, at::Tensor& gradInput2, int pad_size, int kernel_size, int max_displacement,
                              int stride1, int stride2, int /*corr_type_multiply*/) {
  gradInput1.resize_(input1.sizes());
  gradInput2.resize_(input2.sizes());
  auto go = gradOutput.contiguous();
  int st = pcfa_fn2corr_backward(input1.data_ptr<float>(), input2.data_ptr<float>(), go.data_ptr<float>(), gradInput1.data_ptr<float>(),
                                 gradInput2.data_ptr<float>(), (int)input1.size(0), (int)input1.size(1), (int)input1.size(2),
                                 (int)input1.size(3), pad_size, kernel_size, max_displacement, stride1, stride2, cur_stream());
  if (st != PCFA_OK) AT_ERROR("correlation_backward_cuda: ", pcfa_status_string(st));
  return 1;
}

// ------------------------------------------------------------------ resample2d_cuda
int resample2d_cuda_forward(at::Tensor& input1, at::Tensor& input2, at::Tensor& output, int kernel_size, bool bilinear) {
  ok(pcfa_resample2d_forward(input1.data_ptr<float>(), input2.data_ptr<float>(), output.data_ptr<float>(), (int)input1.size(0),
                             (int)input1.size(1), (int)input1.size(2), (int)input1.size(3), (int)input2.size(2), (int)input2.size(3),
                             kernel_size, bilinear ? 1 : 0, cur_stream()), "pcfa_resample2d_forward");
  return 1;
}

int resample2d_cuda_backward(at::Tensor& input1, at::Tensor& input2, at::Tensor& gradOutput, at::Tensor& gradInput1,
                             at::Tensor& gradInput2, int kernel_size, bool bilinear) {
  // gradInput1 must be zero-filled by the caller, as resample2d.py:38 already does
  ok(pcfa_resample2d_backward(input1.data_ptr<float>(), input2.data_ptr<float>(), gradOutput.data_ptr<float>(),
                              gradInput1.data_ptr<float>(), gradInput2.data_ptr<float>(), (int)input1.size(0), (int)input1.size(1),
                              (int)input1.size(2), (int)input1.size(3), (int)input2.size(2), (int)input2.size(3), kernel_size,
                              bilinear ? 1 : 0, cur_stream()), "pcfa_resample2d_backward");
  return 1;
}

// ------------------------------------------------------------------ channelnorm_cuda
int channelnorm_cuda_forward(at::Tensor& input1, at::Tensor& output, int norm_deg) {
  ok(pcfa_channelnorm_forward(input1.data_ptr<float>(), output.data_ptr<float>(), (int)input1.size(0), (int)input1.size(1),
                              (int)input1.size(2), (int)input1.size(3), norm_deg, cur_stream()), "pcfa_channelnorm_forward");
  return 1;
}

int channelnorm_cuda_backward(at::Tensor& input1, at::Tensor& output, at::Tensor& gradOutput, at::Tensor& gradInput1, int norm_deg) {
  ok(pcfa_channelnorm_backward(input1.data_ptr<float>(), output.data_ptr<float>(), gradOutput.data_ptr<float>(),
                               gradInput1.data_ptr<float>(), (int)input1.size(0), (int)input1.size(1), (int)input1.size(2),
                               (int)input1.size(3), norm_deg, cur_stream()), "pcfa_channelnorm_backward");
  return 1;
}

// ------------------------------------------------------------------ CorrBlock (no native module in the reference): the four calls
void corr_block_example(const float* fmap1, const float* fmap2, const float* coords, const float* grad_out, float* pyramid,
                        float* grad_pyramid, float* out, float* grad_fmap1, float* grad_fmap2, void* workspace, int B, int C, int H,
                        int W) {
  int64_t off[5]; int hs[4], ws[4];
  ok(pcfa_corr_pyramid_layout(B, H, W, 4, off, hs, ws), "layout");                           // off[4] floats in total
  const int64_t wsb = pcfa_corr_pyramid_workspace_bytes(B, C, H, W, 4);
  ok(pcfa_corr_pyramid_forward(fmap1, fmap2, pyramid, workspace, wsb, B, C, H, W, 4, /*impl auto*/ 0, cur_stream()), "build");
  ok(pcfa_corr_lookup_forward(pyramid, coords, out, B, H, W, 4, /*radius*/ 4, cur_stream()), "lookup");   // x iterations
  // backward pass: zero grad_pyramid once (cudaMemsetAsync), then per lookup in reverse order:
  ok(pcfa_corr_lookup_backward(grad_out, coords, grad_pyramid, B, H, W, 4, 4, cur_stream()), "lookup backward");
  ok(pcfa_corr_pyramid_backward(grad_pyramid, fmap1, fmap2, grad_fmap1, grad_fmap2, workspace, wsb, B, C, H, W, 4, 0, cur_stream()),
     "build backward");
}

#ifdef PCFA_BUILD_MODULE           // pick one module per build, as the reference's setup.py files do
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward", &correlation_sample_forward, "Spatial Correlation Sampler Forward");
  m.def("backward", &correlation_sample_backward, "Spatial Correlation Sampler backward");
}
#endif
