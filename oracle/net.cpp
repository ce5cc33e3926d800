// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// fp32 CPU restatement of the policy/value network, agent.rs:11-144:
//   input_conv(19->128, 3x3 same) -> input_bn -> relu            agent.rs:114-117
//   10 x ResidualBlock (conv-bn-relu-conv-bn-add-relu)            agent.rs:33-45,118-120
//   policy: conv1x1 128->32, bn, relu, conv1x1 32->64, flatten [64*64], softmax   agent.rs:124-130
//   value : conv1x1 128->8, bn, relu, flatten [8*64], linear 512->64, relu, linear 64->1, tanh   agent.rs:133-141
// Layer semantics (burn 0.18.0 defaults, crate not vendored, [recalled]): conv weight [Cout,Cin,kh,kw] + bias,
// BatchNorm inference (x-mean)/sqrt(var+1e-5)*gamma+beta, Linear weight [d_in,d_out] (y = xW + b).
//
// Array order (the order az_load_weights also uses; names in include/az_b200.h):
//   input_conv.{weight,bias}, input_bn.{gamma,beta,running_mean,running_var},
//   res_blocks.{0..9}.{conv1.{weight,bias}, bn1.{4}, conv2.{weight,bias}, bn2.{4}},
//   policy_conv_1.{weight,bias}, policy_bn.{4}, policy_conv_2.{weight,bias},
//   value_conv.{weight,bias}, value_bn.{4}, value_linear_1.{weight,bias}, value_linear_2.{weight,bias}
#include "oracle.hpp"
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

constexpr int NUM_BLOCKS = 10, FILTERS = 128;
constexpr int N_ARRAYS = 6 + NUM_BLOCKS * 12 + 8 + 10;

struct Net {
    std::vector<std::vector<float>> a;
};

static size_t array_size(int i) {
    auto bn = [](int k, int c) -> size_t { (void)k; return (size_t)c; };
    if (i == 0) return 128 * 19 * 9;
    if (i == 1) return 128;
    if (i < 6) return bn(i, 128);
    i -= 6;
    if (i < NUM_BLOCKS * 12) {
        int j = i % 12;
        if (j == 0 || j == 6) return (size_t)128 * 128 * 9;
        return 128;
    }
    i -= NUM_BLOCKS * 12;
    switch (i) {
        case 0: return 32 * 128; case 1: return 32;
        case 2: case 3: case 4: case 5: return 32;
        case 6: return 64 * 32; case 7: return 64;
        case 8: return 8 * 128; case 9: return 8;
        case 10: case 11: case 12: case 13: return 8;
        case 14: return 512 * 64; case 15: return 64;
        case 16: return 64; case 17: return 1;
    }
    return 0;
}

int net_num_arrays() { return N_ARRAYS; }
size_t net_array_size(int i) { return array_size(i); }
const float* net_array(const Net* n, int i) { return n->a[i].data(); }

Net* net_create_from_arrays(const float* const* arrays, int n_arrays) {
    if (n_arrays != N_ARRAYS) return nullptr;
    Net* n = new Net;
    n->a.resize(N_ARRAYS);
    for (int i = 0; i < N_ARRAYS; i++) n->a[i].assign(arrays[i], arrays[i] + array_size(i));
    return n;
}
Net* net_create_random(u64) { return nullptr; }
void net_destroy(Net* n) { delete n; }

// out[co][64] = bias[co] + sum_k W[co][k] * X[k][64]
static void gemm64(const float* W, const float* bias, const float* X, int cout, int K, float* out) {
    for (int co = 0; co < cout; co++) {
        float acc[64];
        for (int p = 0; p < 64; p++) acc[p] = bias[co];
        const float* w = W + (size_t)co * K;
        for (int k = 0; k < K; k++) {
            float a = w[k];
            const float* x = X + (size_t)k * 64;
            for (int p = 0; p < 64; p++) acc[p] += a * x[p];
        }
        std::memcpy(out + (size_t)co * 64, acc, sizeof acc);
    }
}

// im2col for a 3x3 same-padded conv on an 8x8 board: X[(ci*9 + ky*3 + kx)][r*8+f]
static void im2col3(const float* in, int cin, float* X) {
    for (int ci = 0; ci < cin; ci++)
        for (int ky = 0; ky < 3; ky++)
            for (int kx = 0; kx < 3; kx++) {
                float* x = X + (size_t)(ci * 9 + ky * 3 + kx) * 64;
                for (int r = 0; r < 8; r++)
                    for (int f = 0; f < 8; f++) {
                        int rr = r + ky - 1, ff = f + kx - 1;
                        x[r * 8 + f] = (rr >= 0 && rr < 8 && ff >= 0 && ff < 8) ? in[ci * 64 + rr * 8 + ff] : 0.0f;
                    }
            }
}

static void batchnorm(float* x, int c, const float* gamma, const float* beta, const float* mean, const float* var) {
    for (int ch = 0; ch < c; ch++) {
        float inv = std::sqrt(var[ch] + 1e-5f);
        for (int p = 0; p < 64; p++) x[ch * 64 + p] = (x[ch * 64 + p] - mean[ch]) / inv * gamma[ch] + beta[ch];
    }
}
static void relu(float* x, int n) { for (int i = 0; i < n; i++) x[i] = x[i] > 0.0f ? x[i] : 0.0f; }

void net_forward(const Net* net, const float* planes, float* policy, float* value) {
    const auto& a = net->a;
    std::vector<float> X((size_t)128 * 9 * 64), x(128 * 64), y(128 * 64), z(128 * 64);
    im2col3(planes, 19, X.data());
    gemm64(a[0].data(), a[1].data(), X.data(), 128, 19 * 9, x.data());
    batchnorm(x.data(), 128, a[2].data(), a[3].data(), a[4].data(), a[5].data());
    relu(x.data(), 128 * 64);
    for (int b = 0; b < NUM_BLOCKS; b++) {
        int o = 6 + b * 12;
        im2col3(x.data(), 128, X.data());
        gemm64(a[o].data(), a[o + 1].data(), X.data(), 128, 128 * 9, y.data());
        batchnorm(y.data(), 128, a[o + 2].data(), a[o + 3].data(), a[o + 4].data(), a[o + 5].data());
        relu(y.data(), 128 * 64);
        im2col3(y.data(), 128, X.data());
        gemm64(a[o + 6].data(), a[o + 7].data(), X.data(), 128, 128 * 9, z.data());
        batchnorm(z.data(), 128, a[o + 8].data(), a[o + 9].data(), a[o + 10].data(), a[o + 11].data());
        for (int i = 0; i < 128 * 64; i++) x[i] = z[i] + x[i];
        relu(x.data(), 128 * 64);
    }
    int h = 6 + NUM_BLOCKS * 12;
    // policy head
    std::vector<float> p1(32 * 64), logits(64 * 64);
    gemm64(a[h].data(), a[h + 1].data(), x.data(), 32, 128, p1.data());
    batchnorm(p1.data(), 32, a[h + 2].data(), a[h + 3].data(), a[h + 4].data(), a[h + 5].data());
    relu(p1.data(), 32 * 64);
    gemm64(a[h + 6].data(), a[h + 7].data(), p1.data(), 64, 32, logits.data());
    float mx = logits[0];
    for (int i = 1; i < 4096; i++) mx = logits[i] > mx ? logits[i] : mx;
    double sum = 0.0;
    for (int i = 0; i < 4096; i++) { policy[i] = std::exp(logits[i] - mx); sum += policy[i]; }
    float fs = (float)sum;
    for (int i = 0; i < 4096; i++) policy[i] = policy[i] / fs;
    // value head
    std::vector<float> v1(8 * 64), h1(64);
    gemm64(a[h + 8].data(), a[h + 9].data(), x.data(), 8, 128, v1.data());
    batchnorm(v1.data(), 8, a[h + 10].data(), a[h + 11].data(), a[h + 12].data(), a[h + 13].data());
    relu(v1.data(), 8 * 64);
    const float* W1 = a[h + 14].data();  // [512][64]
    for (int o = 0; o < 64; o++) h1[o] = a[h + 15][o];
    for (int i = 0; i < 512; i++) { float xi = v1[i]; for (int o = 0; o < 64; o++) h1[o] += xi * W1[i * 64 + o]; }
    relu(h1.data(), 64);
    float vl = a[h + 17][0];
    for (int i = 0; i < 64; i++) vl += h1[i] * a[h + 16][i];
    *value = std::tanh(vl);
}

}  // namespace orc
