// CLIP ViT encoders (second compute sink): the arithmetic of the reference's vendored model
// rdm/modules/custom_clip/model.py -- VisualTransformer (:201-235), Transformer / ResidualAttentionBlock (:166-198),
// CLIP.encode_image (:304), CLIP.encode_text (:307-320) -- and of ClipImageRetriever.preprocess (rdm/modules/retrievers.py:83-95).
// Dense layers run on the tcgen05 engine (bf16 hi/lo split by default, gemm_tc.cu); LayerNorm, d_head-64 attention (causal for
// text), embedding / patch / EOT-gather and the bicubic resize are small fp32 kernels.  Parameter names = the reference state dict.
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "../../include/rdm_b200.h"
#include <math_constants.h>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// ---- small kernels -----------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const long long* __restrict__ tok, const float* __restrict__ emb, const float* __restrict__ pos, int T, int W, int vocab,
                                    long long total4, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int V = W / 4; int v = (int)(i % V); long long m = i / V; int t = (int)(m % T);
    long long id = tok[m]; if (id < 0) id = 0; if (id >= vocab) id = vocab - 1;
    float4 a = *reinterpret_cast<const float4*>(emb + id * W + v * 4), p = *reinterpret_cast<const float4*>(pos + (size_t)t * W + v * 4);
    *reinterpret_cast<float4*>(out + m * W + v * 4) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}
// out[b,:] = x[b, argmax_t tok[b,t], :]  (first maximal index, like torch.argmax)
__global__ void gather_eot_kernel(const long long* __restrict__ tok, const float* __restrict__ x, int T, int W, float* __restrict__ out) {
    const int b = blockIdx.x;
    __shared__ int s_t;
    if (threadIdx.x == 0) {
        long long best = tok[(size_t)b * T]; int bt = 0;
        for (int t = 1; t < T; t++) { long long v = tok[(size_t)b * T + t]; if (v > best) { best = v; bt = t; } }
        s_t = bt;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) out[(size_t)b * W + c] = x[((size_t)b * T + s_t) * W + c];
}
__global__ void gather_row0_kernel(const float* __restrict__ x, int T, int W, float* __restrict__ out) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < W; c += blockDim.x) out[(size_t)b * W + c] = x[(size_t)b * T * W + c];
}
// A[(b,gy,gx)][(c,ky,kx)] = img[b,c,gy*P+ky,gx*P+kx]  (im2col of the stride-P patch convolution, K order = conv weight layout)
__global__ void patchify_kernel(const float* __restrict__ img, int R, int P, int G, long long total4, Out4 y) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int K = 3 * P * P, V = K / 4; int v = (int)(i % V); long long m = i / V;
    int gx = (int)(m % G), gy = (int)((m / G) % G), b = (int)(m / ((long long)G * G));
    int k = v * 4, c = k / (P * P), ky = (k / P) % P, kx = k % P;           // P % 4 == 0: the 4 elements share (c, ky)
    const float4 q = *reinterpret_cast<const float4*>(img + (((size_t)b * 3 + c) * R + gy * P + ky) * R + gx * P + kx);
    if (y.f) *reinterpret_cast<float4*>(y.f + m * y.ldf + k) = q;
    if (y.hi) store_planes4(y.hi + m * y.ldb + k, y.lo ? y.lo + m * y.ldb + k : nullptr, y.f16, q.x, q.y, q.z, q.w);
}
// x[b,0,:] = class + pos[0];  x[b,1+g,:] = patch[b*G2+g,:] + pos[1+g]
__global__ void assemble_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos, int T, int W,
                                       long long total4, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int V = W / 4; int v = (int)(i % V); long long m = i / V; int t = (int)(m % T); long long b = m / T;
    float4 a = t == 0 ? *reinterpret_cast<const float4*>(cls + v * 4) : *reinterpret_cast<const float4*>(patch + (b * (T - 1) + t - 1) * W + v * 4);
    float4 p = *reinterpret_cast<const float4*>(pos + (size_t)t * W + v * 4);
    *reinterpret_cast<float4*>(out + m * W + v * 4) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}
// [N, K] -> [K, N]
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    int r = (int)(i / cols), c = (int)(i % cols);
    out[(size_t)c * rows + r] = in[i];
}

// torch upsample_bicubic2d (A = -0.75, align_corners=True, border clamp) fused with (x+1)/2 and the CLIP mean/std normalisation
__device__ __forceinline__ float cub1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cub2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__global__ void clip_preprocess_kernel(const float* __restrict__ in, int B, int H, int W, int S, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * 3 * S * S) return;
    int ox = (int)(i % S), oy = (int)((i / S) % S), c = (int)((i / ((long long)S * S)) % 3), b = (int)(i / ((long long)3 * S * S));
    const float sy = S > 1 ? (float)(H - 1) / (float)(S - 1) : 0.f, sx = S > 1 ? (float)(W - 1) / (float)(S - 1) : 0.f;
    const float ry = sy * oy, rx = sx * ox;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    const float ty = ry - iy, tx = rx - ix;
    const float wy[4] = {cub2(ty + 1.f), cub1(ty), cub1(1.f - ty), cub2(2.f - ty)};
    const float wx[4] = {cub2(tx + 1.f), cub1(tx), cub1(1.f - tx), cub2(2.f - tx)};
    const float* p = in + ((size_t)b * 3 + c) * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; a++) {
        int yy = min(max(iy - 1 + a, 0), H - 1);
        float row = 0.f;
#pragma unroll
        for (int d = 0; d < 4; d++) { int xx = min(max(ix - 1 + d, 0), W - 1); row += p[(size_t)yy * W + xx] * wx[d]; }
        acc += row * wy[a];
    }
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f}, sd[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    out[i] = ((acc + 1.f) * 0.5f - mean[c]) / sd[c];
}

struct Slot { size_t numel = 0; float* dst = nullptr; bool loaded = false; int tr_rows = 0, tr_cols = 0; };   // tr_*: store transposed
struct LayerW { float *ln1g, *ln1b, *inw, *inb, *ow, *ob, *ln2g, *ln2b, *fcw, *fcb, *pw, *pb; };
struct Tower { int W = 0, heads = 0, L = 0, T = 0; std::vector<LayerW> layers; };

struct Opnd {
    View f; __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; int ldb = 0; int f16 = 0;
    bool tc() const { return hi != nullptr; }
    Out4 out4() const { return tc() ? Out4(hi, lo, ldb, f16) : Out4(f); }
};
inline int mode_nsplit(int m) { return m == RDM_UNET_MODE_TC_BF16X3 ? 3 : m == RDM_UNET_MODE_TC_FP16X2 ? 2 : 1; }
inline bool mode_a_split(int m) { return m == RDM_UNET_MODE_TC_BF16X3; }
inline bool mode_w_split(int m) { return m == RDM_UNET_MODE_TC_BF16X3 || m == RDM_UNET_MODE_TC_FP16X2; }
inline int mode_f16(int m) { return (m == RDM_UNET_MODE_TC_FP16X2 || m == RDM_UNET_MODE_TC_FP16) ? 1 : 0; }

}  // namespace

struct rdm_clip {
    int device = 0; rdm_clip_cfg cfg{}; int mode = RDM_UNET_MODE_TC_BF16X3;
    float* wbase = nullptr; size_t wfloats = 0, woff = 0;
    __nv_bfloat16* wb_hi = nullptr; __nv_bfloat16* wb_lo = nullptr; bool planes_dirty = true;
    std::unordered_map<std::string, Slot> params; std::vector<std::string> order;
    Tower vis, txt;
    float *conv1 = nullptr, *cls = nullptr, *vpos = nullptr, *lnpre_g = nullptr, *lnpre_b = nullptr, *lnpost_g = nullptr, *lnpost_b = nullptr, *vprojT = nullptr;
    float *tokemb = nullptr, *tpos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr, *tprojT = nullptr, *logit = nullptr;
    char* ws = nullptr; size_t ws_cap = 0, ws_off = 0;
};

namespace {
typedef rdm_clip Clip;

float* walloc(Clip* n, size_t floats) { floats = (floats + 63) & ~(size_t)63; float* p = n->wbase ? n->wbase + n->woff : (float*)nullptr + n->woff; n->woff += floats; return p; }
float* reg(Clip* n, const std::string& name, size_t numel, int tr_rows = 0, int tr_cols = 0) {
    Slot s; s.numel = numel; s.dst = walloc(n, numel); s.tr_rows = tr_rows; s.tr_cols = tr_cols;
    n->params[name] = s; n->order.push_back(name);
    return s.dst;
}
void build_tower(Clip* n, Tower& t, const std::string& p, int W, int heads, int L, int T) {
    t.W = W; t.heads = heads; t.L = L; t.T = T; t.layers.clear();
    for (int i = 0; i < L; i++) {
        const std::string q = p + "resblocks." + std::to_string(i) + ".";
        LayerW l;
        l.inw = reg(n, q + "attn.in_proj_weight", (size_t)3 * W * W); l.inb = reg(n, q + "attn.in_proj_bias", 3 * W);
        l.ow = reg(n, q + "attn.out_proj.weight", (size_t)W * W); l.ob = reg(n, q + "attn.out_proj.bias", W);
        l.ln1g = reg(n, q + "ln_1.weight", W); l.ln1b = reg(n, q + "ln_1.bias", W);
        l.fcw = reg(n, q + "mlp.c_fc.weight", (size_t)4 * W * W); l.fcb = reg(n, q + "mlp.c_fc.bias", 4 * W);
        l.pw = reg(n, q + "mlp.c_proj.weight", (size_t)4 * W * W); l.pb = reg(n, q + "mlp.c_proj.bias", W);
        l.ln2g = reg(n, q + "ln_2.weight", W); l.ln2b = reg(n, q + "ln_2.bias", W);
        t.layers.push_back(l);
    }
}
void build(Clip* n) {
    const rdm_clip_cfg& c = n->cfg;
    n->woff = 0; n->params.clear(); n->order.clear();
    const int G = c.image_resolution / c.vision_patch_size, vw = c.vision_width, P = c.vision_patch_size, E = c.embed_dim, tw = c.transformer_width;
    n->tpos = reg(n, "positional_embedding", (size_t)c.context_length * tw);
    n->tprojT = reg(n, "text_projection", (size_t)tw * E, tw, E);
    n->logit = reg(n, "logit_scale", 1);
    n->cls = reg(n, "visual.class_embedding", vw);
    n->vpos = reg(n, "visual.positional_embedding", (size_t)(G * G + 1) * vw);
    n->vprojT = reg(n, "visual.proj", (size_t)vw * E, vw, E);
    n->conv1 = reg(n, "visual.conv1.weight", (size_t)vw * 3 * P * P);
    n->lnpre_g = reg(n, "visual.ln_pre.weight", vw); n->lnpre_b = reg(n, "visual.ln_pre.bias", vw);
    build_tower(n, n->vis, "visual.transformer.", vw, vw / 64, c.vision_layers, G * G + 1);
    n->lnpost_g = reg(n, "visual.ln_post.weight", vw); n->lnpost_b = reg(n, "visual.ln_post.bias", vw);
    build_tower(n, n->txt, "transformer.", tw, c.transformer_heads, c.transformer_layers, c.context_length);
    n->tokemb = reg(n, "token_embedding.weight", (size_t)c.vocab_size * tw);
    n->lnf_g = reg(n, "ln_final.weight", tw); n->lnf_b = reg(n, "ln_final.bias", tw);
}

struct Run { Clip* n; cudaStream_t st; int rc = RDM_OK; };
#define RUN(expr) do { if (r.rc == RDM_OK) r.rc = (expr); } while (0)

void* wsalloc(Clip* n, size_t bytes) { bytes = (bytes + 255) & ~(size_t)255; void* p = n->ws + n->ws_off; n->ws_off += bytes; return p; }
View fresh(Clip* n, int M, int C) { return View((float*)wsalloc(n, (size_t)M * C * 4), C, C); }
Opnd fresh_opnd(Clip* n, int M, int C, bool tc) {
    Opnd o; if (!tc) { o.f = fresh(n, M, C); return o; }
    o.hi = (__nv_bfloat16*)wsalloc(n, (size_t)M * C * 2);
    if (mode_a_split(n->mode)) o.lo = (__nv_bfloat16*)wsalloc(n, (size_t)M * C * 2);
    o.ldb = C; o.f.C = C; o.f16 = mode_f16(n->mode); return o;
}
Opnd from_view(View v) { Opnd o; o.f = v; return o; }
bool tc_ok(Clip* n, int K) { return n->mode != RDM_UNET_MODE_FP32 && K % 64 == 0; }

// out = epi(a [M,K] * w[N,K]^T)
void gemm(Run& r, const Opnd& a, int M, int K, const float* w, const float* bias, int N, GemmEpi e, const Opnd& out) {
    Clip* n = r.n;
    e.bias = bias;
    if (a.tc()) {
        TcA ta; ta.hi = a.hi; ta.lo = a.lo; ta.ld = a.ldb; ta.B = M; ta.H = 1; ta.W = 1; ta.C = K; ta.ksize = 1;
        TcW tw; tw.hi = n->wb_hi + (w - n->wbase); tw.lo = mode_w_split(n->mode) ? n->wb_lo + (w - n->wbase) : nullptr; tw.N = N; tw.K = K; tw.ld = K;
        const int ns = mode_nsplit(n->mode), f16 = mode_f16(n->mode);
        if (out.tc()) { e.out = nullptr; RUN(gemm_tc(ta, tw, e, out.hi, out.lo, out.ldb, ns, f16, r.st)); }
        else { e.out = out.f.p; e.out_ld = out.f.ld; RUN(gemm_tc(ta, tw, e, nullptr, nullptr, 0, ns, f16, r.st)); }
        return;
    }
    GemmA ga; ga.x = a.f.p; ga.ld = a.f.ld; ga.B = M; ga.Cin = K;
    if (out.tc()) {
        View tmp = fresh(n, M, N); e.out = tmp.p; e.out_ld = tmp.ld;
        RUN(gemm_simt(ga, w, N, e, r.st));
        RUN(k_split_planes(tmp, M, out.out4(), r.st));
    } else { e.out = out.f.p; e.out_ld = out.f.ld; RUN(gemm_simt(ga, w, N, e, r.st)); }
}

// x: [B*T, W] fp32 -> returns the output view (fresh buffers from the workspace)
View run_tower(Run& r, const Tower& t, View x, int B, int causal) {
    Clip* n = r.n; const int M = B * t.T, W = t.W;
    const bool tc = tc_ok(n, W);
    for (const LayerW& l : t.layers) {
        Opnd nrm = fresh_opnd(n, M, W, tc);
        RUN(k_layernorm(x, M, l.ln1g, l.ln1b, 1e-5f, nrm.out4(), r.st));
        View qkv = fresh(n, M, 3 * W);
        gemm(r, nrm, M, W, l.inw, l.inb, 3 * W, GemmEpi(), from_view(qkv));
        Opnd att = fresh_opnd(n, M, W, tc);
        RUN(k_attention_d64(qkv.cols(0, W), qkv.cols(W, W), qkv.cols(2 * W, W), B, t.T, t.heads, 0.125f, causal, att.out4(), r.st));
        View x2 = fresh(n, M, W);
        { GemmEpi e; e.res = x.p; e.res_ld = x.ld; gemm(r, att, M, W, l.ow, l.ob, W, e, from_view(x2)); }
        Opnd n2 = fresh_opnd(n, M, W, tc);
        RUN(k_layernorm(x2, M, l.ln2g, l.ln2b, 1e-5f, n2.out4(), r.st));
        Opnd h = fresh_opnd(n, M, 4 * W, tc);
        { GemmEpi e; e.act = ACT_QUICKGELU; gemm(r, n2, M, W, l.fcw, l.fcb, 4 * W, e, h); }
        View x3 = fresh(n, M, W);
        { GemmEpi e; e.res = x2.p; e.res_ld = x2.ld; gemm(r, h, M, 4 * W, l.pw, l.pb, W, e, from_view(x3)); }
        x = x3;
    }
    return x;
}

int ensure_ws(Clip* n, size_t bytes) {
    n->ws_off = 0;
    if (bytes <= n->ws_cap) return RDM_OK;
    if (n->ws) cudaFree(n->ws);
    n->ws = nullptr; n->ws_cap = 0;
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->ws, bytes));
    n->ws_cap = bytes;
    return RDM_OK;
}
size_t tower_bytes(const Tower& t, int B) {           // generous upper bound of the per-call workspace
    size_t M = (size_t)B * t.T;
    return t.L * (M * t.W * 4 * 14 + 8192) + M * t.W * 64 + ((size_t)1 << 20);
}
int ensure_planes(Clip* n, cudaStream_t st) {
    if (n->mode == RDM_UNET_MODE_FP32 || !n->planes_dirty) return RDM_OK;
    if (!n->wb_hi) RDM_CHECK_CUDA(cudaMalloc((void**)&n->wb_hi, n->wfloats * 2));
    if (!n->wb_lo) RDM_CHECK_CUDA(cudaMalloc((void**)&n->wb_lo, n->wfloats * 2));
    RDM_TRY(k_split_planes(View(n->wbase, 64, 64), (long long)(n->wfloats / 64), Out4(n->wb_hi, n->wb_lo, 64, mode_f16(n->mode)), st));
    n->planes_dirty = false;
    return RDM_OK;
}
int64_t missing(const Clip* n) { int64_t m = 0; for (auto& kv : n->params) if (!kv.second.loaded && kv.first != "logit_scale") m++; return m; }
inline int blocks_for(long long n, int t) { return (int)((n + t - 1) / t); }
#define LAUNCH_CHECK() do { RDM_COUNT_LAUNCH(); RDM_CHECK_CUDA(cudaGetLastError()); } while (0)

}  // namespace

extern "C" {

int rdm_clip_create(rdm_clip_t** out, const rdm_clip_cfg* c, int32_t device) {
    RDM_REQUIRE(out && c, RDM_ERR_ARG, "rdm_clip_create: null argument");
    RDM_REQUIRE(c->vision_width % 64 == 0 && c->transformer_width == c->transformer_heads * 64, RDM_ERR_UNSUPPORTED,
                "rdm_clip_create: only d_head = 64 towers are implemented (vision_width %d, text %d/%d)", c->vision_width, c->transformer_width, c->transformer_heads);
    RDM_REQUIRE(c->vision_patch_size % 4 == 0 && c->image_resolution % c->vision_patch_size == 0 && c->embed_dim % 4 == 0, RDM_ERR_UNSUPPORTED, "rdm_clip_create: patch/resolution");
    DeviceGuard guard(device);
    RDM_REQUIRE(guard.ok, RDM_ERR_CUDA, "rdm_clip_create: cannot select device %d", device);
    rdm_clip* n = new rdm_clip(); n->device = device; n->cfg = *c;
    build(n); n->wfloats = n->woff;
    if (cudaMalloc((void**)&n->wbase, n->wfloats * 4) != cudaSuccess) { delete n; rdm_set_error("rdm_clip_create: cudaMalloc failed"); return RDM_ERR_CUDA; }
    cudaMemset(n->wbase, 0, n->wfloats * 4);
    build(n);
    *out = n; return RDM_OK;
}
void rdm_clip_destroy(rdm_clip_t* n) {
    if (!n) return;
    DeviceGuard guard(n->device);
    for (void* p : {(void*)n->wbase, (void*)n->wb_hi, (void*)n->wb_lo, (void*)n->ws}) if (p) cudaFree(p);
    delete n;
}
int64_t rdm_clip_num_params(const rdm_clip_t* n) { return n ? (int64_t)n->order.size() : 0; }
const char* rdm_clip_param_name(const rdm_clip_t* n, int64_t i) { return (n && i >= 0 && i < (int64_t)n->order.size()) ? n->order[i].c_str() : nullptr; }
int64_t rdm_clip_param_numel(const rdm_clip_t* n, const char* name) { if (!n || !name) return -1; auto it = n->params.find(name); return it == n->params.end() ? -1 : (int64_t)it->second.numel; }
int64_t rdm_clip_missing(const rdm_clip_t* n) { return n ? missing(n) : -1; }
int rdm_clip_set_mode(rdm_clip_t* n, int32_t mode) {
    RDM_REQUIRE(n && mode >= 0 && mode <= RDM_UNET_MODE_TC_FP16, RDM_ERR_ARG, "rdm_clip_set_mode: bad argument");
    if (mode != n->mode) n->planes_dirty = true;
    n->mode = mode; return RDM_OK;
}
int rdm_clip_load(rdm_clip_t* n, const char* name, const float* host, int64_t numel) {
    RDM_REQUIRE(n && name && host, RDM_ERR_ARG, "rdm_clip_load: null argument");
    auto it = n->params.find(name);
    RDM_REQUIRE(it != n->params.end(), RDM_ERR_ARG, "rdm_clip_load: unknown parameter '%s'", name);
    Slot& s = it->second;
    RDM_REQUIRE((size_t)numel == s.numel, RDM_ERR_ARG, "rdm_clip_load: '%s' has %lld elements, expected %zu", name, (long long)numel, s.numel);
    DeviceGuard guard(n->device);
    if (s.tr_rows) {        // [rows, cols] -> [cols, rows]  (projection matrices become K-major)
        std::vector<float> t((size_t)numel);
        for (int r = 0; r < s.tr_rows; r++) for (int c = 0; c < s.tr_cols; c++) t[(size_t)c * s.tr_rows + r] = host[(size_t)r * s.tr_cols + c];
        RDM_CHECK_CUDA(cudaMemcpy(s.dst, t.data(), (size_t)numel * 4, cudaMemcpyHostToDevice));
    } else RDM_CHECK_CUDA(cudaMemcpy(s.dst, host, (size_t)numel * 4, cudaMemcpyHostToDevice));
    s.loaded = true; n->planes_dirty = true;
    return RDM_OK;
}

int rdm_clip_encode_text(rdm_clip_t* n, const int64_t* tokens, int32_t B, float* out, void* stream) {
    RDM_REQUIRE(n && tokens && out && B >= 1, RDM_ERR_ARG, "rdm_clip_encode_text: bad argument");
    RDM_REQUIRE(missing(n) == 0, RDM_ERR_STATE, "rdm_clip_encode_text: %lld parameters not loaded", (long long)missing(n));
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Tower& t = n->txt; const int T = t.T, W = t.W, E = n->cfg.embed_dim;
    RDM_TRY(ensure_ws(n, tower_bytes(t, B)));
    RDM_TRY(ensure_planes(n, st));
    Run r{n, st};
    View x = fresh(n, B * T, W);
    const long long total4 = (long long)B * T * (W / 4);
    embed_tokens_kernel<<<blocks_for(total4, 256), 256, 0, st>>>((const long long*)tokens, n->tokemb, n->tpos, T, W, n->cfg.vocab_size, total4, x.p);
    LAUNCH_CHECK();
    x = run_tower(r, t, x, B, 1);
    View xe = fresh(n, B, W);
    gather_eot_kernel<<<B, 128, 0, st>>>((const long long*)tokens, x.p, T, W, xe.p);
    LAUNCH_CHECK();
    Opnd nf = fresh_opnd(n, B, W, tc_ok(n, W));
    RUN(k_layernorm(xe, B, n->lnf_g, n->lnf_b, 1e-5f, nf.out4(), st));
    gemm(r, nf, B, W, n->tprojT, nullptr, E, GemmEpi(), from_view(View(out, E, E)));
    return r.rc;
}

int rdm_clip_encode_image(rdm_clip_t* n, const float* img, int32_t B, float* out, void* stream) {
    RDM_REQUIRE(n && img && out && B >= 1, RDM_ERR_ARG, "rdm_clip_encode_image: bad argument");
    RDM_REQUIRE(missing(n) == 0, RDM_ERR_STATE, "rdm_clip_encode_image: %lld parameters not loaded", (long long)missing(n));
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    const Tower& t = n->vis; const int T = t.T, W = t.W, E = n->cfg.embed_dim, P = n->cfg.vision_patch_size, R = n->cfg.image_resolution, G = R / P, K = 3 * P * P;
    RDM_TRY(ensure_ws(n, tower_bytes(t, B) + (size_t)B * G * G * K * 12));
    RDM_TRY(ensure_planes(n, st));
    Run r{n, st};
    Opnd col = fresh_opnd(n, B * G * G, K, tc_ok(n, K));
    const long long t4 = (long long)B * G * G * (K / 4);
    patchify_kernel<<<blocks_for(t4, 256), 256, 0, st>>>(img, R, P, G, t4, col.out4());
    LAUNCH_CHECK();
    View pe = fresh(n, B * G * G, W);
    gemm(r, col, B * G * G, K, n->conv1, nullptr, W, GemmEpi(), from_view(pe));
    View x0 = fresh(n, B * T, W), x = fresh(n, B * T, W);
    const long long a4 = (long long)B * T * (W / 4);
    assemble_tokens_kernel<<<blocks_for(a4, 256), 256, 0, st>>>(pe.p, n->cls, n->vpos, T, W, a4, x0.p);
    LAUNCH_CHECK();
    RUN(k_layernorm(x0, B * T, n->lnpre_g, n->lnpre_b, 1e-5f, Out4(x), st));
    x = run_tower(r, t, x, B, 0);
    View x_cls = fresh(n, B, W);
    gather_row0_kernel<<<B, 128, 0, st>>>(x.p, T, W, x_cls.p);
    LAUNCH_CHECK();
    Opnd nf = fresh_opnd(n, B, W, tc_ok(n, W));
    RUN(k_layernorm(x_cls, B, n->lnpost_g, n->lnpost_b, 1e-5f, nf.out4(), st));
    gemm(r, nf, B, W, n->vprojT, nullptr, E, GemmEpi(), from_view(View(out, E, E)));
    return r.rc;
}

int rdm_clip_preprocess(const float* img, int32_t B, int32_t H, int32_t W, int32_t size, float* out, int32_t device, void* stream) {
    RDM_REQUIRE(img && out && B >= 1 && H >= 1 && W >= 1 && size >= 1, RDM_ERR_ARG, "rdm_clip_preprocess: bad argument");
    DeviceGuard guard(device);
    const long long n = (long long)B * 3 * size * size;
    clip_preprocess_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(img, B, H, W, size, out);
    LAUNCH_CHECK();
    return RDM_OK;
}

}  // extern "C"
