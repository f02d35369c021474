// Fused multi-tensor SGD step for the reference's PolyOptimizer (tool/torchutils.py:11-33, used at contrast_train.py:91-96,
// 397-399): the step after the hot path.  One launch updates up to kSgdMaxTensors parameter tensors:
//   d = grad + wd * p;   buf = first ? d : momentum * buf + (1 - dampening) * d;   p -= lr * buf      (torch.optim.SGD)
// The hyper-parameter quirks of the reference (its weight decay lands in SGD's momentum slot) are reproduced on the host
// side (wseg_b200/optim.py); this kernel is plain SGD.
#include <algorithm>

#include "common.cuh"

namespace wsegk {

constexpr int kSgdMaxTensors = 48;
struct SgdTable {
    float* p[kSgdMaxTensors];
    const float* g[kSgdMaxTensors];
    float* m[kSgdMaxTensors];
    long long n[kSgdMaxTensors];
};

__global__ void __launch_bounds__(256) sgd_multi_kernel(SgdTable t, float lr, float wd, float mom, float damp, int first) {
    const int k = blockIdx.y;
    float* __restrict__ p = t.p[k];
    const float* __restrict__ g = t.g[k];
    float* __restrict__ m = t.m[k];
    const long long n = t.n[k];
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float pv = p[i];
        float d = __fmaf_rn(wd, pv, g[i]);
        if (mom != 0.f) {
            const float b = first ? d : __fmaf_rn(mom, m[i], (1.f - damp) * d);
            m[i] = b;
            d = b;
        }
        p[i] = __fmaf_rn(-lr, d, pv);
    }
}

}  // namespace wsegk

using namespace wsegk;

extern "C" int wsegk_sgd_step(float* const* params, const float* const* grads, float* const* momentum_bufs,
                              const long long* numels, int n_tensors, float lr, float weight_decay, float momentum,
                              float dampening, int first_step, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(params && grads && numels && n_tensors > 0, WSEGK_ERR_NULL, "sgd_step: null table / no tensors");
    WSEGK_REQUIRE(momentum == 0.f || momentum_bufs != nullptr, WSEGK_ERR_NULL, "sgd_step: momentum needs buffers");
    int launches = 0;
    for (int base = 0; base < n_tensors; base += kSgdMaxTensors) {
        SgdTable t{};
        const int cnt = std::min(kSgdMaxTensors, n_tensors - base);
        long long longest = 0;
        for (int k = 0; k < cnt; ++k) {
            WSEGK_REQUIRE(params[base + k] && grads[base + k] && numels[base + k] > 0, WSEGK_ERR_NULL,
                          "sgd_step: tensor %d has a null pointer or no elements", base + k);
            t.p[k] = params[base + k];
            t.g[k] = grads[base + k];
            t.m[k] = momentum_bufs ? momentum_bufs[base + k] : nullptr;
            WSEGK_REQUIRE(momentum == 0.f || t.m[k] != nullptr, WSEGK_ERR_NULL, "sgd_step: tensor %d has no momentum buffer",
                          base + k);
            t.n[k] = numels[base + k];
            longest = std::max(longest, t.n[k]);
        }
        const int bx = static_cast<int>(std::min<long long>((longest + 255) / 256, 64));
        sgd_multi_kernel<<<dim3(bx, cnt), 256, 0, stream>>>(t, lr, weight_decay, momentum, dampening, first_step);
        ++launches;
    }
    count_launches(launches);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}
