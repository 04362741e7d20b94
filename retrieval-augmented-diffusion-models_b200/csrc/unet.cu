// U-Net executor: builds the layer plan of rdm/modules/diffusionmodules/openaimodel.py:66-317 from a config
// struct, owns the (re-packed) weights, the workspace arena and the per-context cross-attention K/V, and
// runs UNetModel.forward (openaimodel.py:335-371) as a fixed sequence of kernels on one stream.
//
// Data layout: NHWC fp32 activations as [M, C] matrices (kernels.cuh).  Skip connections are written by their
// producer straight into the concat buffer of the consuming output block (torch.cat at openaimodel.py:365 is free).
// Hoisted out of the step loop: cross-attention K/V projections of the (step-invariant) retrieved context
// (attention.py:47-48 -> rdm_unet_set_context), and all 25 ResBlock emb_layers as ONE GEMM per forward.
#include "kernels.cuh"
#include "gemm_tc.cuh"
#include "../../include/rdm_b200.h"
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>
#include <memory>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {

struct Arena {
    char* base = nullptr; size_t cap = 0, off = 0, peak = 0; bool dry = true;
    float* allocf(size_t n) { return (float*)alloc(n * sizeof(float)); }
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        size_t o = off; off += bytes; if (off > peak) peak = off;
        return dry ? (void*)((char*)256 + o) : (void*)(base + o);
    }
    size_t mark() const { return off; }
    void release(size_t m) { off = m; }
};

struct ParamSlot { size_t numel = 0; bool loaded = false; std::function<void(const float*, std::vector<float>&)> pack; float* dst = nullptr; size_t dst_numel = 0;
                   // generic scatter: packed data is copied to dst (dst_numel floats) -- or rows scattered (see rows_*).
                   int rows = 0, row_len = 0, dst_row0 = 0, dst_row_step = 1; };

struct Norm { int C = 0; float* g = nullptr; float* b = nullptr; };
struct Conv { int cin = 0, cout = 0, ks = 1; float* w = nullptr; float* b = nullptr; float* wfold = nullptr; };     // w: [cout][ks*ks*cin] (tap-major K); wfold: [4][cout][2*2*cin], conv after Upsample only (fold_up_weights)
struct Lin { int in = 0, out = 0; float* w = nullptr; float* b = nullptr; };                // w: [out][in]

struct ResW { int cin, cout; Norm n1; Conv c1; int emb_off; Norm n2; Conv c2; bool has_skip; Conv skip; float eps = 1e-5f; float* b2s = nullptr; };   // b2s: c2.b + skip.b (ensure_weight_planes)   // emb_off < 0: no time embedding (VQ decoder ResnetBlock)
struct AttnW { int C; Norm norm; Lin qkv; Conv proj_out; };     // taming/ldm AttnBlock: single head over all pixels, q/k/v 1x1 convs concatenated [3C, C]
enum DecKind { D_RES, D_ATTN, D_UP };
struct DecLayer { DecKind kind; int idx; };
struct STW { int C, heads, id; Norm norm; Conv proj_in; Norm ln1, ln2, ln3; Lin qkv, o1, q2, kv2, o2, ff1, ff2; Conv proj_out; };
enum LayerKind { L_CONV_IN, L_RES, L_ST, L_DOWN, L_UP };
struct Layer { LayerKind kind; int idx; };
struct Block { std::vector<Layer> layers; int cout; int ds_after; };     // ds_after: spatial divisor after this block

struct Act { View v; int B, H, W; int M() const { return B * H * W; } };

}  // namespace

struct rdm_unet {
    int device = 0;
    rdm_unet_cfg cfg{};
    int mode = 0;
    // weights
    float* wbase = nullptr; size_t wfloats = 0, woff = 0;
    __nv_bfloat16* wb_hi = nullptr; __nv_bfloat16* wb_lo = nullptr; bool planes_dirty = true;   // bf16 planes of the weight arena (same offsets)
    std::unordered_map<std::string, ParamSlot> params;
    std::vector<std::string> param_order;
    std::vector<ResW> res; std::vector<STW> sts; std::vector<Conv> convs;   // convs: conv_in, downs, ups
    std::vector<Block> in_blocks, out_blocks; Block mid;
    Lin te0, te2, emb_all; Norm out_norm; Conv out_conv;
    int emb_total = 0, ted = 0;
    std::vector<int> skip_ch;          // channels of hs[i]
    // first-stage VQ decoder (kind == 1): ldm.modules.diffusionmodules.model.Decoder behind VQModelInterface.decode
    int kind = 0; rdm_vqdec_cfg dcfg{};
    std::vector<AttnW> attns; std::vector<DecLayer> dec_layers; Conv dec_conv_in, dec_conv_out; Norm dec_norm_out;
    float* codebook = nullptr; float* pq_w = nullptr; float* pq_b = nullptr;
    cudaGraphExec_t dec_exec = nullptr; int dec_key[5] = {0, 0, 0, 0, -1}; unsigned long long dec_kernels = 0; float* dec_in = nullptr; float* dec_out = nullptr; size_t dec_io_cap = 0;
    // runtime
    Arena arena; double* stats = nullptr; size_t stats_cap = 0, stats_off = 0;
    // batch chains: the B2 rows of a forward are split into `chains` sub-batches that run the whole layer sequence concurrently on their own
    // streams (fork / join by events, also inside the captured graphs); every chain owns a slice of the arena and of the statistics buffer
    int chains = getenv("RDM_CHAINS") ? atoi(getenv("RDM_CHAINS")) : 1; int plan_chains = 0;
    size_t shared_bytes = 0, chain_bytes = 0, chain_stats = 0;
    cudaStream_t side[8] = {nullptr}; cudaEvent_t ev_fork = nullptr, ev_join[8] = {nullptr};
    float* ctx_kv = nullptr; size_t ctx_kv_floats = 0; std::vector<size_t> ctx_off; int ctx_B = 0, ctx_k = 0;
    long long* t_dev = nullptr; int t_cap = 0;
    int plan_B = 0, plan_H = 0, plan_W = 0;
    std::vector<float*> host_scratch;
    int debug = 0, debug_block = 0; std::string debug_log;
    // CUDA-graph replay: fixed staging buffers + one captured graph per (Bx, B2, H, W, mode)
    int use_graph = 1;
    float* x_in = nullptr; long long* t_in = nullptr; float* eps_buf = nullptr; float* x_state = nullptr; float* p0_buf = nullptr; int* step_dev = nullptr;
    size_t io_cap = 0, tin_cap = 0;         // capacities of the staging buffers: elements of x_in / eps_buf / ..., entries of t_in
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t fwd_exec = nullptr; int fwd_key[6] = {0, 0, 0, 0, -1, 0};
    int skip = getenv("RDM_SKIP") ? atoi(getenv("RDM_SKIP")) : 0;      // ablation mask (see RUN_UNLESS)
    cudaGraphExec_t step_exec = nullptr; long long step_key[10] = {0};
    unsigned long long fwd_kernels = 0, step_kernels = 0;      // kernels inside each captured graph (for rdm_launch_count)
    // profiling (rdm_unet_profile_forward): event pairs around every GEMM launch
    int profile = 0; std::vector<cudaEvent_t> prof_ev; std::vector<double> prof_flops; std::vector<int> prof_kind;
    std::vector<std::string> prof_desc; std::string prof_text;
};

namespace {

typedef rdm_unet Net;

float* walloc(Net* n, size_t floats) {
    floats = (floats + 63) & ~(size_t)63;
    float* p = n->wbase ? n->wbase + n->woff : (float*)nullptr + n->woff;   // pass 1 counts, pass 2 places
    n->woff += floats;
    return p;
}

// ---- parameter registration (names follow SURVEY.md Appendix C) ---------------------------------------
void reg_plain(Net* n, const std::string& name, float* dst, size_t numel) {
    ParamSlot s; s.numel = numel; s.dst = dst; s.dst_numel = numel;
    s.pack = [](const float* src, std::vector<float>& out) { (void)src; (void)out; };
    n->params[name] = s; n->param_order.push_back(name);
}
// conv weight [cout, cin, ks, ks] -> [cout][ks*ks][cin]
void reg_conv_w(Net* n, const std::string& name, float* dst, int cout, int cin, int ks) {
    ParamSlot s; s.numel = (size_t)cout * cin * ks * ks; s.dst = dst; s.dst_numel = s.numel;
    s.pack = [=](const float* src, std::vector<float>& out) {
        out.resize((size_t)cout * cin * ks * ks);
        const int T = ks * ks;
        for (int o = 0; o < cout; o++)
            for (int c = 0; c < cin; c++)
                for (int t = 0; t < T; t++) out[((size_t)o * T + t) * cin + c] = src[((size_t)o * cin + c) * T + t];
    };
    n->params[name] = s; n->param_order.push_back(name);
}
// rows of a [rows, row_len] matrix scattered to dst rows dst_row0 + r*step (QKV concat, GEGLU interleave, emb concat)
void reg_rows(Net* n, const std::string& name, float* dst_base, int rows, int row_len, int dst_row0, int step) {
    ParamSlot s; s.numel = (size_t)rows * row_len; s.dst = dst_base; s.rows = rows; s.row_len = row_len; s.dst_row0 = dst_row0; s.dst_row_step = step;
    n->params[name] = s; n->param_order.push_back(name);
}
// GEGLU proj [8C, C]: value rows j -> 2j, gate rows 4C+j -> 2j+1
void reg_geglu(Net* n, const std::string& name, float* dst, int half_rows, int row_len) {
    ParamSlot s; s.numel = (size_t)2 * half_rows * row_len; s.dst = dst; s.dst_numel = s.numel;
    s.pack = [=](const float* src, std::vector<float>& out) {
        out.resize((size_t)2 * half_rows * row_len);
        for (int j = 0; j < half_rows; j++) {
            memcpy(&out[(size_t)(2 * j) * row_len], &src[(size_t)j * row_len], row_len * sizeof(float));
            memcpy(&out[(size_t)(2 * j + 1) * row_len], &src[(size_t)(half_rows + j) * row_len], row_len * sizeof(float));
        }
    };
    n->params[name] = s; n->param_order.push_back(name);
}

Norm make_norm(Net* n, const std::string& p, int C) {
    Norm r; r.C = C; r.g = walloc(n, C); r.b = walloc(n, C);
    reg_plain(n, p + ".weight", r.g, C); reg_plain(n, p + ".bias", r.b, C);
    return r;
}
Conv make_conv(Net* n, const std::string& p, int cin, int cout, int ks) {
    Conv c; c.cin = cin; c.cout = cout; c.ks = ks; c.w = walloc(n, (size_t)cout * cin * ks * ks); c.b = walloc(n, cout);
    reg_conv_w(n, p + ".weight", c.w, cout, cin, ks); reg_plain(n, p + ".bias", c.b, cout);
    return c;
}
// the 3x3 conv that follows a 2x nearest Upsample: also reserves the parity-folded weights (filled by ensure_weight_planes)
Conv make_up_conv(Net* n, const std::string& p, int cin, int cout) {
    Conv c = make_conv(n, p, cin, cout, 3);
    c.wfold = walloc(n, (size_t)16 * cout * cin);
    return c;
}
Lin make_lin(Net* n, const std::string& p, int in, int out, bool bias) {
    Lin l; l.in = in; l.out = out; l.w = walloc(n, (size_t)in * out); reg_plain(n, p + ".weight", l.w, (size_t)in * out);
    if (bias) { l.b = walloc(n, out); reg_plain(n, p + ".bias", l.b, out); }
    return l;
}

int add_res(Net* n, const std::string& p, int cin, int cout) {
    ResW r{}; r.cin = cin; r.cout = cout;
    r.n1 = make_norm(n, p + ".in_layers.0", cin);
    r.c1 = make_conv(n, p + ".in_layers.2", cin, cout, 3);
    r.emb_off = n->emb_total; n->emb_total += cout;          // emb_layers.1 registered after emb_all is allocated
    r.n2 = make_norm(n, p + ".out_layers.0", cout);
    r.c2 = make_conv(n, p + ".out_layers.3", cout, cout, 3);
    r.has_skip = cin != cout;
    if (r.has_skip) { r.skip = make_conv(n, p + ".skip_connection", cin, cout, 1); r.b2s = walloc(n, cout); }
    n->res.push_back(r);
    n->host_scratch.push_back(nullptr);
    return (int)n->res.size() - 1;
}
std::vector<std::string> g_res_prefix;   // parallel to net->res during construction (single-threaded build)

int add_st(Net* n, const std::string& p, int C, int heads, int ctx_dim) {
    STW s{}; s.C = C; s.heads = heads; s.id = (int)n->sts.size();
    s.norm = make_norm(n, p + ".norm", C);
    s.proj_in = make_conv(n, p + ".proj_in", C, C, 1);
    const std::string t = p + ".transformer_blocks.0";
    s.qkv.in = C; s.qkv.out = 3 * C; s.qkv.w = walloc(n, (size_t)3 * C * C);
    reg_rows(n, t + ".attn1.to_q.weight", s.qkv.w, C, C, 0, 1);
    reg_rows(n, t + ".attn1.to_k.weight", s.qkv.w, C, C, C, 1);
    reg_rows(n, t + ".attn1.to_v.weight", s.qkv.w, C, C, 2 * C, 1);
    s.o1 = make_lin(n, t + ".attn1.to_out.0", C, C, true);
    s.q2 = make_lin(n, t + ".attn2.to_q", C, C, false);
    s.kv2.in = ctx_dim; s.kv2.out = 2 * C; s.kv2.w = walloc(n, (size_t)2 * C * ctx_dim);
    reg_rows(n, t + ".attn2.to_k.weight", s.kv2.w, C, ctx_dim, 0, 1);
    reg_rows(n, t + ".attn2.to_v.weight", s.kv2.w, C, ctx_dim, C, 1);
    s.o2 = make_lin(n, t + ".attn2.to_out.0", C, C, true);
    s.ff1.in = C; s.ff1.out = 8 * C; s.ff1.w = walloc(n, (size_t)8 * C * C); s.ff1.b = walloc(n, 8 * C);
    reg_geglu(n, t + ".ff.net.0.proj.weight", s.ff1.w, 4 * C, C);
    reg_geglu(n, t + ".ff.net.0.proj.bias", s.ff1.b, 4 * C, 1);
    s.ff2 = make_lin(n, t + ".ff.net.2", 4 * C, C, true);
    s.ln1 = make_norm(n, t + ".norm1", C); s.ln2 = make_norm(n, t + ".norm2", C); s.ln3 = make_norm(n, t + ".norm3", C);
    s.proj_out = make_conv(n, p + ".proj_out", C, C, 1);
    n->sts.push_back(s);
    return s.id;
}

bool in_list(const int* v, int n, int x) { for (int i = 0; i < n; i++) if (v[i] == x) return true; return false; }

// Mirrors UNetModel.__init__ (openaimodel.py:137-311).  Called twice: counting pass (wbase == nullptr) and placing pass.
void build_net(Net* n) {
    const rdm_unet_cfg& c = n->cfg;
    n->woff = 0; n->params.clear(); n->param_order.clear(); n->res.clear(); n->sts.clear(); n->convs.clear();
    n->in_blocks.clear(); n->out_blocks.clear(); n->mid = Block(); n->emb_total = 0; n->skip_ch.clear(); g_res_prefix.clear();
    const int mc = c.model_channels; n->ted = 4 * mc;
    n->te0 = make_lin(n, "time_embed.0", mc, n->ted, true);
    n->te2 = make_lin(n, "time_embed.2", n->ted, n->ted, true);
    auto heads_of = [&](int ch) { return c.num_head_channels > 0 ? ch / c.num_head_channels : c.num_heads; };
    auto push_res = [&](Block& b, const std::string& p, int cin, int cout) { b.layers.push_back({L_RES, add_res(n, p, cin, cout)}); g_res_prefix.push_back(p); };

    int ch = mc, ds = 1;
    { Block b; n->convs.push_back(make_conv(n, "input_blocks.0.0", c.in_channels, mc, 3)); b.layers.push_back({L_CONV_IN, 0}); b.cout = mc; b.ds_after = 1; n->in_blocks.push_back(b); n->skip_ch.push_back(mc); }
    for (int level = 0; level < c.n_channel_mult; level++) {
        int mult = c.channel_mult[level];
        for (int r = 0; r < c.num_res_blocks; r++) {
            Block b; std::string p = "input_blocks." + std::to_string(n->in_blocks.size());
            push_res(b, p + ".0", ch, mult * mc); ch = mult * mc;
            if (in_list(c.attention_resolutions, c.n_attention_resolutions, ds)) b.layers.push_back({L_ST, add_st(n, p + ".1", ch, heads_of(ch), c.context_dim)});
            b.cout = ch; b.ds_after = ds; n->in_blocks.push_back(b); n->skip_ch.push_back(ch);
        }
        if (level != c.n_channel_mult - 1) {
            Block b; std::string p = "input_blocks." + std::to_string(n->in_blocks.size());
            n->convs.push_back(make_conv(n, p + ".0.op", ch, ch, 3)); b.layers.push_back({L_DOWN, (int)n->convs.size() - 1});
            ds *= 2; b.cout = ch; b.ds_after = ds; n->in_blocks.push_back(b); n->skip_ch.push_back(ch);
        }
    }
    { Block& b = n->mid; push_res(b, "middle_block.0", ch, ch); b.layers.push_back({L_ST, add_st(n, "middle_block.1", ch, heads_of(ch), c.context_dim)});
      push_res(b, "middle_block.2", ch, ch); b.cout = ch; b.ds_after = ds; }
    std::vector<int> chans = n->skip_ch;
    for (int level = c.n_channel_mult - 1; level >= 0; level--) {
        int mult = c.channel_mult[level];
        for (int i = 0; i <= c.num_res_blocks; i++) {
            int ich = chans.back(); chans.pop_back();
            Block b; std::string p = "output_blocks." + std::to_string(n->out_blocks.size());
            int li = 0;
            push_res(b, p + "." + std::to_string(li++), ch + ich, mc * mult); ch = mc * mult;
            if (in_list(c.attention_resolutions, c.n_attention_resolutions, ds)) b.layers.push_back({L_ST, add_st(n, p + "." + std::to_string(li++), ch, heads_of(ch), c.context_dim)});
            if (level && i == c.num_res_blocks) {
                n->convs.push_back(make_up_conv(n, p + "." + std::to_string(li++) + ".conv", ch, ch)); b.layers.push_back({L_UP, (int)n->convs.size() - 1});
                ds /= 2;
            }
            b.cout = ch; b.ds_after = ds; n->out_blocks.push_back(b);
        }
    }
    n->out_norm = make_norm(n, "out.0", ch);
    n->out_conv = make_conv(n, "out.2", mc, c.out_channels, 3);
    // all ResBlock emb_layers.1 as one [sum(cout), ted] matrix
    n->emb_all.in = n->ted; n->emb_all.out = n->emb_total;
    n->emb_all.w = walloc(n, (size_t)n->emb_total * n->ted); n->emb_all.b = walloc(n, n->emb_total);
    for (size_t i = 0; i < n->res.size(); i++) {
        reg_rows(n, g_res_prefix[i] + ".emb_layers.1.weight", n->emb_all.w, n->res[i].cout, n->ted, n->res[i].emb_off, 1);
        reg_rows(n, g_res_prefix[i] + ".emb_layers.1.bias", n->emb_all.b, n->res[i].cout, 1, n->res[i].emb_off, 1);
    }
}

// ---- forward ------------------------------------------------------------------------------------------
struct Ctx {
    Net* n; cudaStream_t st; bool dry; int rc = RDM_OK; Arena* A = nullptr; size_t stats_off = 0; int b0 = 0;     // A / stats_off / b0: this chain's arena, statistics slice and first batch row
    // column ranges of statistics slabs (View::st) that a GEMM epilogue of THIS forward has filled: (first entry, columns)
    std::vector<std::pair<const double*, int>> st_done;
    void mark_stats(const View& v) { if (v.st) st_done.push_back({v.st, v.C}); }
    bool has_stats(const View& v) const {                 // every column of the view is covered (a concat buffer has two producers)
        if (!v.st) return false;
        const double* cur = v.st; const double* end = v.st + 2 * (size_t)v.C;
        while (cur < end) {
            bool found = false;
            for (const auto& r : st_done) if (r.first == cur) { cur += 2 * (size_t)r.second; found = true; break; }
            if (!found) return false;
        }
        return cur == end;
    }
};
#define RUN(expr) do { if (!cx.dry && cx.rc == RDM_OK) cx.rc = (expr); } while (0)
// Timing ablation (rdm_unet_set_ablation, tools/ablate_forward.py; env RDM_SKIP sets the initial value): a bit mask of kernel classes that are NOT launched (results are garbage; only the
// change of the graph-replayed forward time is meaningful).  1 gn_stats, 2 gn_apply, 4 layernorm, 8 attention, 16 GEMM M>=8192, 32 GEMM M<8192.
#define RUN_UNLESS(bit, expr) do { if (!(cx.n->skip & (bit))) RUN(expr); } while (0)

// GEMM operand / result: an fp32 view (CUDA-core engine) or bf16 hi/lo planes (tcgen05 engine)
struct Opnd {
    View f; __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; int ldb = 0;
    bool tc() const { return hi != nullptr; }
    int f16 = 0;
    Out4 out4() const { return tc() ? Out4(hi, lo, ldb, f16) : Out4(f); }
};
Opnd from_view(View v) { Opnd o; o.f = v; return o; }
// engine configuration of a mode: MMAs per product, whether activations / weights carry a lo plane, fp16 vs bf16 planes
inline int mode_nsplit(int m) { return m == RDM_UNET_MODE_TC_BF16X3 ? 3 : m == RDM_UNET_MODE_TC_FP16X2 ? 2 : 1; }
inline bool mode_a_split(int m) { return m == RDM_UNET_MODE_TC_BF16X3; }
inline bool mode_w_split(int m) { return m == RDM_UNET_MODE_TC_BF16X3 || m == RDM_UNET_MODE_TC_FP16X2; }
inline int mode_f16(int m) { return (m == RDM_UNET_MODE_TC_FP16X2 || m == RDM_UNET_MODE_TC_FP16) ? 1 : 0; }

double* stats_alloc(Ctx& cx, int B, int groups) {
    Net* n = cx.n; size_t need = (size_t)B * groups * 2;
    size_t o = cx.stats_off; cx.stats_off += need;
    return cx.dry ? (double*)nullptr + o : n->stats + o;
}
View fresh(Ctx& cx, int M, int C) { return View(cx.A->allocf((size_t)M * C), C, C); }
// statistics slab [B][C] x {sum, sumsq} for a buffer whose consumer is a GroupNorm (tensor-core modes; zeroed with the rest of n->stats)
void attach_stats(Ctx& cx, View& v, int B) {
    if (cx.n->mode == RDM_UNET_MODE_FP32 || cx.n->kind != 0) return;
    // A/B (RDM_GN_EPI_STATS=1): per-(image, channel) sums accumulated by the epilogue of the producing tcgen05 GEMM (gemm_tc.cu: epi_stats)
    // and folded by the one-read form of gn_fused.cu.  Measured on B200 (full architecture, B2 = 32, fp16): it applies to 12 of the 67
    // GroupNorms of a forward (non-split-K producers at the 32x32 / 16x16 levels whose every input half has such a producer) and is a
    // wash -- 4.645 ms with, 4.635 ms without: the epilogue's shuffles + shared atomics + two named barriers per tile and the larger
    // statistics memset cost what the 12 saved statistics launches gain.  Default off.
    static const int on = getenv("RDM_GN_EPI_STATS") ? atoi(getenv("RDM_GN_EPI_STATS")) : 0;
    if (!on) return;
    const size_t need = (size_t)B * v.C * 2, o = cx.stats_off; cx.stats_off += need;
    v.st = cx.dry ? (double*)nullptr + o : cx.n->stats + o; v.st_ld = v.C;
}
View fresh_st(Ctx& cx, int B, int M, int C) { View v = fresh(cx, M, C); attach_stats(cx, v, B); return v; }
Opnd fresh_opnd(Ctx& cx, int M, int C, bool tc) {
    Opnd o;
    if (!tc) { o.f = fresh(cx, M, C); return o; }
    o.hi = (__nv_bfloat16*)cx.A->alloc((size_t)M * C * 2);
    if (mode_a_split(cx.n->mode)) o.lo = (__nv_bfloat16*)cx.A->alloc((size_t)M * C * 2);
    o.ldb = C; o.f.C = C; o.f16 = mode_f16(cx.n->mode);
    return o;
}
bool tc_ok(Ctx& cx, int B, int H, int W, int C, int ks) {
    if (cx.n->mode == RDM_UNET_MODE_FP32) return false;
    TcA a; a.B = B; a.H = H; a.W = W; a.C = C; a.ksize = ks; a.ld = C;
    return gemm_tc_supported(a);
}
const __nv_bfloat16* w_hi(Net* n, const float* w) { return n->wb_hi + (w - n->wbase); }
const __nv_bfloat16* w_lo(Net* n, const float* w) { return mode_w_split(n->mode) ? n->wb_lo + (w - n->wbase) : nullptr; }

// out = epi(conv/linear(a)).  a: [B*H*W, C] operand; ks 1|3 (stride 1, pad ks/2) on the tensor-core engine;
// stride / ups only exist on the CUDA-core engine (the TC path materialises im2col / upsampled planes instead).
// K extension of a tensor-core GEMM (gemm_tc.cuh: TcA::hi2): a second [M, C] operand with its own [N, C] weights in the same accumulator
struct KExt { const Opnd* a = nullptr; const float* w = nullptr; int C = 0; };
void gemm_any(Ctx& cx, const Opnd& a, int B, int H, int W, int C, int ks, int stride, int ups, const float* w, const float* bias, int N,
              GemmEpi e, const Opnd& out, KExt x2 = KExt()) {
    Net* n = cx.n;
    if (!e.bias) e.bias = bias;
    const int Ho = ups ? H * 2 : (stride == 2 ? (H + 1) / 2 : H), Wo = ups ? W * 2 : (stride == 2 ? (W + 1) / 2 : W);
    const int M = B * Ho * Wo, Nout = e.act == ACT_GEGLU ? N / 2 : N;
    struct ProfScope {
        Ctx& cx; bool on;
        ProfScope(Ctx& c, double flops, int kind) : cx(c), on(!c.dry && c.n->profile) {
            if (!on) return;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cx.n->prof_ev.push_back(e0); cx.n->prof_ev.push_back(e1); cx.n->prof_flops.push_back(flops); cx.n->prof_kind.push_back(kind);
            rec(e0);
        }
        // profile == 2: the forward is being captured into a CUDA graph; external event-record nodes time the kernels inside the replay
        void rec(cudaEvent_t e) { if (cx.n->profile == 2) cudaEventRecordWithFlags(e, cx.st, cudaEventRecordExternal); else cudaEventRecord(e, cx.st); }
        ~ProfScope() { if (on) rec(cx.n->prof_ev.back()); }
    } prof(cx, 2.0 * M * (double)N * (ks * ks * C + x2.C), a.tc() ? 1 : 0);       // ALGORITHMIC flops (the reference's conv); an upsample-folded conv executes 4/9 of them
    if (prof.on) { char d[128]; snprintf(d, sizeof(d), "%s M=%d N=%d K=%d ks=%d HxW=%dx%d act=%d%s%s", a.tc() ? "tc" : "simt", M, N, ks * ks * C + x2.C, ks, H, W, e.act, a.tc() && ups ? " upfold" : "", x2.a ? " kext" : ""); n->prof_desc.push_back(d); }
    if (a.tc()) {
        TcA ta; ta.hi = a.hi; ta.lo = a.lo; ta.ld = a.ldb; ta.B = B; ta.H = H; ta.W = W; ta.C = C; ta.ksize = ks; ta.ups = ups;
        TcW tw; tw.hi = w_hi(n, w); tw.lo = w_lo(n, w); tw.N = N; tw.K = (ups ? 4 : ks * ks) * C; tw.ld = tw.K;      // (ups: `w` = the folded weights)
        if (x2.a) { ta.hi2 = x2.a->hi; ta.ld2 = x2.a->ldb; ta.C2 = x2.C; tw.hi2 = w_hi(n, x2.w); tw.ld2 = x2.C; }
        const int nsplit = mode_nsplit(n->mode), f16 = mode_f16(n->mode);
        const int skip_bit = M >= 8192 ? 16 : 32;
        if (out.tc()) { e.out = nullptr; RUN_UNLESS(skip_bit, gemm_tc(ta, tw, e, out.hi, out.lo, out.ldb, nsplit, f16, cx.st)); }
        else {
            int fused = 0;
            if (out.f.st && B > 0 && M % B == 0) { e.stats = out.f.st; e.stats_ld = out.f.st_ld; e.stats_hw = M / B; e.stats_fused = &fused; }
            e.out = out.f.p; e.out_ld = out.f.ld; RUN_UNLESS(skip_bit, gemm_tc(ta, tw, e, nullptr, nullptr, 0, nsplit, f16, cx.st));
            if (fused && cx.rc == RDM_OK) { View done = out.f; done.C = Nout; cx.mark_stats(done); }
        }
        return;
    }
    GemmA ga; ga.x = a.f.p; ga.ld = a.f.ld; ga.B = B; ga.Hs = H; ga.Ws = W; ga.Cin = C; ga.ksize = ks; ga.stride = stride; ga.ups = ups; ga.Ho = Ho; ga.Wo = Wo;
    if (out.tc()) {      // CUDA-core producer feeding a tensor-core consumer: fp32 temporary, then split
        size_t mk = cx.A->mark();
        View tmp = fresh(cx, M, Nout);
        e.out = tmp.p; e.out_ld = tmp.ld;
        RUN(gemm_simt(ga, w, N, e, cx.st));
        RUN(k_split_planes(tmp, M, out.out4(), cx.st));
        cx.A->release(mk);
    } else {
        e.out = out.f.p; e.out_ld = out.f.ld;
        RUN(gemm_simt(ga, w, N, e, cx.st));
    }
}
void conv_any(Ctx& cx, const Opnd& a, const Act& g, const Conv& c, GemmEpi e, const Opnd& out, KExt x2 = KExt()) {
    gemm_any(cx, a, g.B, g.H, g.W, c.cin, c.ks, 1, 0, c.w, c.b, c.cout, e, out, x2);
}
void lin_any(Ctx& cx, const Opnd& a, int M, const Lin& l, GemmEpi e, const Opnd& out) {
    gemm_any(cx, a, M, 1, 1, l.in, 1, 1, 0, l.w, l.b, l.out, e, out);
}

void gn(Ctx& cx, const Act& x, const Norm& nm, float eps, int silu, const Opnd& y, const Opnd* raw = nullptr) {
    // Statistics and normalisation in ONE launch (gn_fused.cu: the CTAs of an image exchange partial sums through distributed shared
    // memory) for SMALL images -- at the 8x8 / 4x4 levels the two kernels below are pure launch latency (4.32 vs 4.38 ms per forward on
    // B200).  For the large images the one-launch form loses (4.85 vs 4.60 ms when used everywhere: an 8-CTA cluster per image offers
    // the memory system far fewer independent loads than the 148 x 16 small CTAs of gn_stats / gn_apply), so they keep the two kernels.
    // RDM_GN_FUSED_MAX_HW: largest image (pixels) that takes the one-launch form (0: never).
    static const int fused_max_hw = getenv("RDM_GN_FUSED_MAX_HW") ? atoi(getenv("RDM_GN_FUSED_MAX_HW")) : 64;
    if (cx.n->mode != RDM_UNET_MODE_FP32 && x.H * x.W <= fused_max_hw && k_gn_fused_supported(x.v.C, x.H * x.W, 32, false)) {
        RUN_UNLESS(2, k_gn_fused(x.v, x.B, x.H * x.W, 32, nullptr, 0, eps, nm.g, nm.b, silu, y.out4(), raw ? raw->out4() : Out4(), cx.st));
        return;
    }
    if (!cx.dry && cx.has_stats(x.v) && k_gn_fused_supported(x.v.C, x.H * x.W, 32, true)) {
        // the producing GEMMs left per-(image, channel) sums behind (gemm_tc.cu: epi_stats): no statistics pass, x is read once
        RUN_UNLESS(2, k_gn_fused(x.v, x.B, x.H * x.W, 32, x.v.st, x.v.st_ld, eps, nm.g, nm.b, silu, y.out4(), raw ? raw->out4() : Out4(), cx.st));
        return;
    }
    double* s = stats_alloc(cx, x.B, 32);
    RUN_UNLESS(1, k_gn_stats(x.v, x.B, x.H * x.W, 32, s, cx.st));
    RUN_UNLESS(2, k_gn_apply(x.v, x.B, x.H * x.W, 32, s, eps, nm.g, nm.b, silu, y.out4(), raw ? raw->out4() : Out4(), cx.st));
}

void run_res(Ctx& cx, const ResW& r, const Act& x, const float* emb_all, View out) {
    Arena& A = *cx.A; size_t mk = A.mark();
    const int M = x.M();
    const bool tc1 = tc_ok(cx, x.B, x.H, x.W, r.cin, 3), tc2 = tc_ok(cx, x.B, x.H, x.W, r.cout, 3);
    const bool tcs = r.has_skip && tc_ok(cx, x.B, x.H, x.W, r.cin, 1);
    Opnd a1 = fresh_opnd(cx, M, r.cin, tc1), xraw;
    if (tcs) xraw = fresh_opnd(cx, M, r.cin, true);
    gn(cx, x, r.n1, r.eps, 1, a1, tcs ? &xraw : nullptr);
    View h1 = fresh_st(cx, x.B, M, r.cout);                      // consumed by out_layers' GroupNorm only: conv1's epilogue leaves its statistics
    { GemmEpi e; if (r.emb_off >= 0) { e.rowvec = emb_all + r.emb_off; e.rowvec_ld = cx.n->emb_total; e.rows_per_batch = x.H * x.W; }
      conv_any(cx, a1, x, r.c1, e, from_view(h1)); }
    Opnd a2 = fresh_opnd(cx, M, r.cout, tc2);
    gn(cx, Act{h1, x.B, x.H, x.W}, r.n2, r.eps, 1, a2);
    // One-plane tensor-core modes: the 1x1 skip_connection is a K extension of the second conv -- its k-blocks (raw fp16 x, skip weights)
    // follow the nine taps into the same TMEM accumulator, the summed bias replaces both biases; no second GEMM, no fp32 residual buffer
    // written and read back.  RDM_RES_SKIP_FUSED=0: the separate GEMM (also what the split-plane and fp32 modes run).
    static const int fuse_on = getenv("RDM_RES_SKIP_FUSED") ? atoi(getenv("RDM_RES_SKIP_FUSED")) : 1;
    if (fuse_on && tcs && tc2 && r.b2s && mode_nsplit(cx.n->mode) == 1) {
        GemmEpi e; e.bias = r.b2s;
        KExt x2; x2.a = &xraw; x2.w = r.skip.w; x2.C = r.cin;
        conv_any(cx, a2, x, r.c2, e, from_view(out), x2);
        A.release(mk);
        return;
    }
    View resv = x.v;
    if (r.has_skip) { resv = fresh(cx, M, r.cout); conv_any(cx, tcs ? xraw : from_view(x.v), x, r.skip, GemmEpi(), from_view(resv)); }
    { GemmEpi e; e.res = resv.p; e.res_ld = resv.ld; conv_any(cx, a2, x, r.c2, e, from_view(out)); }
    A.release(mk);
}

void run_st(Ctx& cx, const STW& s, const Act& x, View out) {
    Net* n = cx.n; Arena& A = *cx.A; size_t mk = A.mark();
    const int M = x.M(), C = s.C, N = x.H * x.W;
    const float scale = 0.17677669529663687f;                      // d_head ** -0.5 for d_head = 32 (attention.py:27)
    const bool tcp = tc_ok(cx, M, 1, 1, C, 1), tcf = tc_ok(cx, M, 1, 1, 4 * C, 1);
    Opnd a = fresh_opnd(cx, M, C, tcp);
    gn(cx, x, s.norm, 1e-6f, 0, a);
    View t0 = fresh(cx, M, C);
    conv_any(cx, a, x, s.proj_in, GemmEpi(), from_view(t0));
    Opnd nrm = fresh_opnd(cx, M, C, tcp);
    // self-attention (attention.py:93)
    RUN_UNLESS(4, k_layernorm(t0, M, s.ln1.g, s.ln1.b, 1e-5f, nrm.out4(), cx.st));
    View qkv = fresh(cx, M, 3 * C);
    Opnd att = fresh_opnd(cx, M, C, tcp);
    if (tcp && mode_f16(n->mode) && k_attention_mma_supported(N, N)) {
        // f16 engine modes: the QKV GEMM writes ONE fp16 plane that the warp-MMA attention kernel consumes directly
        size_t mk2 = A.mark();
        Opnd qh; qh.hi = (__nv_bfloat16*)A.alloc((size_t)M * 3 * C * 2); qh.ldb = 3 * C; qh.f.C = 3 * C; qh.f16 = 1;
        lin_any(cx, nrm, M, s.qkv, GemmEpi(), qh);
        const __half* qp = reinterpret_cast<const __half*>(qh.hi);
        RUN_UNLESS(8, k_attention_mma(qp, qp + C, qp + 2 * C, 3 * C, x.B, N, N, s.heads, scale, att.out4(), cx.st));
        A.release(mk2);
    } else {
        lin_any(cx, nrm, M, s.qkv, GemmEpi(), from_view(qkv));
        RUN_UNLESS(8, k_attention(qkv.cols(0, C), qkv.cols(C, C), qkv.cols(2 * C, C), x.B, N, N, s.heads, scale, att.out4(), cx.st));
    }
    View t1 = fresh(cx, M, C);
    { GemmEpi e; e.res = t0.p; e.res_ld = t0.ld; lin_any(cx, att, M, s.o1, e, from_view(t1)); }
    // cross-attention to the retrieved neighbours (attention.py:94); K/V were projected once in set_context
    RUN_UNLESS(4, k_layernorm(t1, M, s.ln2.g, s.ln2.b, 1e-5f, nrm.out4(), cx.st));
    View kv(cx.dry ? nullptr : n->ctx_kv + n->ctx_off[s.id] + (size_t)cx.b0 * n->ctx_k * 2 * C, 2 * C, 2 * C);     // this chain's rows of the projected context
    if (tcp && n->ctx_k <= 8 && N >= 16 && N % 16 == 0 && !(n->skip & 64)) {
        // tensor-core engine: softmax(q k^T) v over the k retrieved neighbours runs in the epilogue of the to_q GEMM (one 32-column
        // accumulator chunk = one head's query), so neither q nor a separate attention launch exists
        GemmEpi e; e.act = ACT_XATTN; e.xkv = kv.p; e.xkv_ld = kv.ld; e.xv_off = C; e.xk = n->ctx_k; e.xscale = scale; e.rows_per_batch = N;
        gemm_any(cx, nrm, M, 1, 1, s.q2.in, 1, 1, 0, s.q2.w, nullptr, s.q2.out, e, att);
    } else {
        View q2 = qkv.cols(0, C);                                    // reuse the qkv buffer
        lin_any(cx, nrm, M, s.q2, GemmEpi(), from_view(q2));
        RUN_UNLESS(8, k_attention(q2, kv.cols(0, C), kv.cols(C, C), x.B, N, n->ctx_k, s.heads, scale, att.out4(), cx.st));
    }
    View t2 = t0;                                                // t0 is dead after t1 was formed
    { GemmEpi e; e.res = t1.p; e.res_ld = t1.ld; lin_any(cx, att, M, s.o2, e, from_view(t2)); }
    // GEGLU feed-forward (attention.py:95)
    RUN_UNLESS(4, k_layernorm(t2, M, s.ln3.g, s.ln3.b, 1e-5f, nrm.out4(), cx.st));
    Opnd g = fresh_opnd(cx, M, 4 * C, tcf);
    { GemmEpi e; e.act = ACT_GEGLU; lin_any(cx, nrm, M, s.ff1, e, g); }
    Opnd t3 = tcp ? fresh_opnd(cx, M, C, true) : from_view(t1);  // only proj_out consumes t3
    { GemmEpi e; e.res = t2.p; e.res_ld = t2.ld; lin_any(cx, g, M, s.ff2, e, t3); }
    { GemmEpi e; e.res = x.v.p; e.res_ld = x.v.ld; conv_any(cx, t3, x, s.proj_out, e, from_view(out)); }
    A.release(mk);
}

void run_down(Ctx& cx, const Conv& c, const Act& x, View out) {
    Arena& A = *cx.A; size_t mk = A.mark();
    const int Ho = (x.H + 1) / 2, Wo = (x.W + 1) / 2, Mo = x.B * Ho * Wo;
    if (tc_ok(cx, Mo, 1, 1, 9 * c.cin, 1)) {
        Opnd col = fresh_opnd(cx, Mo, 9 * c.cin, true);
        RUN(k_im2col_s2(x.v, x.B, x.H, x.W, col.out4(), cx.st));
        gemm_any(cx, col, Mo, 1, 1, 9 * c.cin, 1, 1, 0, c.w, c.b, c.cout, GemmEpi(), from_view(out));
    } else {
        gemm_any(cx, from_view(x.v), x.B, x.H, x.W, c.cin, 3, 2, 0, c.w, c.b, c.cout, GemmEpi(), from_view(out));
    }
    A.release(mk);
}
void run_up(Ctx& cx, const Conv& c, const Act& x, View out) {
    Arena& A = *cx.A; size_t mk = A.mark();
    // Tensor-core modes: the upsampled plane is never built -- each output parity is a 2x2-tap conv of the SOURCE image with pre-summed
    // weights (gemm_tc.cuh: TcA::ups): 4/9 of the MMA work, and a same-size fp16 copy of x instead of a 4x larger one.  The folded weights
    // are rounded to the operand format after the fp32 sums (so a product differs from the reference's by rounding only).
    // RDM_UP_FOLD=0: upsample2x + plain 3x3 conv.
    static const int fold_on = getenv("RDM_UP_FOLD") ? atoi(getenv("RDM_UP_FOLD")) : 1;
    if (fold_on && c.wfold && c.cout % 32 == 0 && x.W <= 128 && tc_ok(cx, x.B, x.H, x.W, c.cin, 3)) {
        Opnd src = fresh_opnd(cx, x.M(), c.cin, true);
        RUN(k_split_planes(x.v, x.M(), src.out4(), cx.st));
        gemm_any(cx, src, x.B, x.H, x.W, c.cin, 3, 1, 1, c.wfold, c.b, c.cout, GemmEpi(), from_view(out));
        A.release(mk);
        return;
    }
    if (tc_ok(cx, x.B, 2 * x.H, 2 * x.W, c.cin, 3)) {
        Opnd up = fresh_opnd(cx, x.B * 4 * x.H * x.W, c.cin, true);
        RUN(k_upsample2x(x.v, x.B, x.H, x.W, up.out4(), cx.st));
        gemm_any(cx, up, x.B, 2 * x.H, 2 * x.W, c.cin, 3, 1, 0, c.w, c.b, c.cout, GemmEpi(), from_view(out));
    } else {
        gemm_any(cx, from_view(x.v), x.B, x.H, x.W, c.cin, 3, 1, 1, c.w, c.b, c.cout, GemmEpi(), from_view(out));
    }
    A.release(mk);
}

// debug: "<tag> mean absmean" of an [M, C] view (synchronises; only when rdm_unet_set_debug(h, 1))
void debug_view(Ctx& cx, const char* tag, int a, int b, View v, int M) {
    if (cx.dry || !cx.n->debug || cx.rc != RDM_OK) return;
    std::vector<float> host((size_t)M * v.C);
    cudaError_t err = cudaStreamSynchronize(cx.st);
    if (err != cudaSuccess) { cx.n->debug_log += std::string("sync failed: ") + cudaGetErrorString(err) + "\n"; return; }
    if (cudaMemcpy2D(host.data(), (size_t)v.C * 4, v.p, (size_t)v.ld * 4, (size_t)v.C * 4, M, cudaMemcpyDeviceToHost) != cudaSuccess) { cx.n->debug_log += "memcpy failed\n"; return; }
    double s = 0, sa = 0; for (float f : host) { s += f; sa += f < 0 ? -f : f; }
    char line[160]; snprintf(line, sizeof(line), "%s %d %d %.9g %.9g\n", tag, a, b, s / host.size(), sa / host.size());
    cx.n->debug_log += line;
}

// Runs the layers of one block; the LAST layer writes into `dst`.
Act run_block(Ctx& cx, const Block& b, Act x, const float* emb_all, View dst) {
    Net* n = cx.n;
    for (size_t i = 0; i < b.layers.size(); i++) {
        const Layer& L = b.layers[i];
        const bool last = i + 1 == b.layers.size();
        int Ho = x.H, Wo = x.W, Co = 0;
        switch (L.kind) {
            case L_CONV_IN: Co = n->convs[L.idx].cout; break;
            case L_RES: Co = n->res[L.idx].cout; break;
            case L_ST: Co = n->sts[L.idx].C; break;
            case L_DOWN: Co = n->convs[L.idx].cout; Ho = (x.H + 1) / 2; Wo = (x.W + 1) / 2; break;
            case L_UP: Co = n->convs[L.idx].cout; Ho = x.H * 2; Wo = x.W * 2; break;
        }
        const bool feeds_gn = !last && (b.layers[i + 1].kind == L_RES || b.layers[i + 1].kind == L_ST);
        View o = last ? dst : feeds_gn ? fresh_st(cx, x.B, x.B * Ho * Wo, Co) : fresh(cx, x.B * Ho * Wo, Co);
        switch (L.kind) {
            case L_CONV_IN: {
                const Conv& c = n->convs[L.idx];
                if (k_conv_first_supported(c.cin, c.cout)) RUN(k_conv_first(x.v, x.B, x.H, x.W, c.w, c.b, c.cout, o, cx.st));
                else gemm_any(cx, from_view(x.v), x.B, x.H, x.W, c.cin, 3, 1, 0, c.w, c.b, c.cout, GemmEpi(), from_view(o));
                break;
            }
            case L_RES: run_res(cx, n->res[L.idx], x, emb_all, o); break;
            case L_ST: run_st(cx, n->sts[L.idx], x, o); break;
            case L_DOWN: run_down(cx, n->convs[L.idx], x, o); break;
            case L_UP: run_up(cx, n->convs[L.idx], x, o); break;
        }
        x = Act{View(o.p, o.ld, Co), x.B, Ho, Wo};
        debug_view(cx, "layer", cx.n->debug_block, (int)i, x.v, x.M());
    }
    cx.n->debug_block++;
    return x;
}

// One chain: the whole layer sequence for batch rows [cx.b0, cx.b0 + nb) of the forward.  x_nchw / t / emb_all / out_nchw are the FULL-batch
// buffers; this chain reads and writes only its rows.
int chain_impl(Ctx& cx, const float* x_nchw, int Bx, const long long* t, int nb, int H, int W, View emb_all, float* out_nchw) {
    Net* n = cx.n; Arena& A = *cx.A; cudaStream_t st = cx.st; const bool dry = cx.dry;
    const rdm_unet_cfg& c = n->cfg;
    const int nin = (int)n->in_blocks.size(), nout = (int)n->out_blocks.size();
    const float* emb = dry ? emb_all.p : emb_all.p + (size_t)cx.b0 * emb_all.ld;
    std::vector<int> hH(nin), hW(nin);
    { int h = H, w = W; for (int i = 0; i < nin; i++) { if (n->in_blocks[i].layers[0].kind == L_DOWN) { h = (h + 1) / 2; w = (w + 1) / 2; } hH[i] = h; hW[i] = w; } }
    // concat buffers: output block j reads cat([h (ch_j), hs[nin-1-j] (ich_j)])
    std::vector<View> cat(nout); std::vector<int> cat_ch(nout);
    {
        int ch = n->mid.cout;
        for (int j = 0; j < nout; j++) {
            int i = nin - 1 - j, ich = n->skip_ch[i];
            cat_ch[j] = ch;
            cat[j] = View(A.allocf((size_t)nb * hH[i] * hW[i] * (ch + ich)), ch + ich, ch + ich);
            attach_stats(cx, cat[j], nb);                        // both halves are GroupNorm inputs (the skip half also of the next input block)
            ch = n->out_blocks[j].cout;
        }
    }
    View x0 = fresh(cx, nb * H * W, c.in_channels < 4 ? 4 : c.in_channels); x0.C = c.in_channels;
    RUN(k_nchw_to_nhwc(x_nchw, Bx, nb, c.in_channels, H, W, x0, st, cx.b0));
    Act h{x0, nb, H, W};
    for (int i = 0; i < nin; i++) {
        int j = nout - 1 - i;
        h = run_block(cx, n->in_blocks[i], h, emb, cat[j].cols(cat_ch[j], n->skip_ch[i]));
    }
    h = run_block(cx, n->mid, h, emb, cat[0].cols(0, cat_ch[0]));
    View last;
    for (int j = 0; j < nout; j++) {
        int i = nin - 1 - j;
        Act in{cat[j], nb, hH[i], hW[i]};
        View dst;
        if (j + 1 < nout) dst = cat[j + 1].cols(0, cat_ch[j + 1]);
        else { last = fresh_st(cx, nb, nb * H * W, n->out_blocks[j].cout); dst = last; }
        h = run_block(cx, n->out_blocks[j], in, emb, dst);
    }
    // out = conv3x3(SiLU(GN(h)))  (openaimodel.py:312-316,371)
    Opnd a = fresh_opnd(cx, h.M(), h.v.C, tc_ok(cx, h.B, h.H, h.W, h.v.C, 3));
    gn(cx, h, n->out_norm, 1e-5f, 1, a);
    View o = fresh(cx, h.M(), 4); o.C = c.out_channels;
    conv_any(cx, a, h, n->out_conv, GemmEpi(), from_view(o));
    debug_view(cx, "out", 0, 0, o, h.M());
    RUN(k_nhwc_to_nchw(o, nb, c.out_channels, H, W, dry ? out_nchw : out_nchw + (size_t)cx.b0 * c.out_channels * H * W, st));
    return cx.rc;
}

int effective_chains(const Net* n, int B2) {
    int g = n->chains < 1 ? 1 : n->chains > 8 ? 8 : n->chains;
    if (n->debug || n->profile) g = 1;                  // per-layer dumps / per-GEMM event brackets follow ONE stream
    return g < B2 ? g : B2;
}

// x_nchw: [Bx, C, H, W] with Bx == B2 or B2/2 (then duplicated, ddim.py:233); t: int64 [B2] device; out NCHW [B2,...]
int forward_impl(Net* n, const float* x_nchw, int Bx, const long long* t, int B2, int H, int W, float* out_nchw, cudaStream_t st, bool dry) {
    const int G = dry ? (n->plan_chains > 0 ? n->plan_chains : 1) : n->plan_chains;
    Ctx cx{n, st, dry};
    n->debug_block = 0; if (!dry) n->debug_log.clear();
    Arena& A = n->arena; A.off = 0; A.dry = dry; n->stats_off = 0;
    cx.A = &A;
    const rdm_unet_cfg& c = n->cfg;
    if (!dry) RDM_CHECK_CUDA(cudaMemsetAsync(n->stats, 0, n->stats_cap * sizeof(double), st));
    // time embedding (openaimodel.py:352-353); only SiLU(emb) is ever consumed (ResBlock emb_layers = [SiLU, Linear]).
    // M = B2 rows: pure weight streaming (38 MB for the 25 concatenated emb_layers) -- on the tensor-core engine the weights arrive by
    // TMA and split-K spreads them over all SMs; in fp32 mode the CUDA-core engine is used.  Shared by all chains (runs before the fork).
    View temb = fresh(cx, B2, c.model_channels), semb = fresh(cx, B2, n->ted), emb_all = fresh(cx, B2, n->emb_total);
    RUN(k_timestep_embedding(t, B2, c.model_channels, temb.p, st));
    {
        const bool tc0 = tc_ok(cx, B2, 1, 1, c.model_channels, 1), tc1 = tc_ok(cx, B2, 1, 1, n->ted, 1);
        Opnd t0 = from_view(temb);
        if (tc0) { t0 = fresh_opnd(cx, B2, c.model_channels, true); RUN(k_split_planes(temb, B2, t0.out4(), st)); }
        Opnd e1 = fresh_opnd(cx, B2, n->ted, tc1), se = fresh_opnd(cx, B2, n->ted, tc1);
        { GemmEpi e; e.act = ACT_SILU; lin_any(cx, t0, B2, n->te0, e, e1); }
        { GemmEpi e; e.act = ACT_SILU; lin_any(cx, e1, B2, n->te2, e, se); }
        lin_any(cx, se, B2, n->emb_all, GemmEpi(), from_view(emb_all));
        if (cx.n->debug && !dry) { if (se.tc()) { /* planes only: no fp32 copy to show */ } else semb = se.f; }
    }
    debug_view(cx, "temb", 0, 0, temb, B2); debug_view(cx, "emb_all", 0, 0, emb_all, B2);
    if (cx.rc != RDM_OK) return cx.rc;
    if (dry) {
        // plan: shared prologue, then the LARGEST chain (the first B2 % G chains carry one more row); every chain gets a slice of that size
        n->shared_bytes = A.off;
        const int nbmax = (B2 + G - 1) / G;
        Arena ca; ca.dry = true;
        Ctx cc{n, st, true}; cc.A = &ca;
        RDM_TRY(chain_impl(cc, x_nchw, Bx, t, nbmax, H, W, emb_all, out_nchw));
        n->chain_bytes = ca.peak; n->chain_stats = cc.stats_off;
        A.peak = n->shared_bytes + (size_t)G * n->chain_bytes;
        n->stats_off = (size_t)G * n->chain_stats;
        return RDM_OK;
    }
    // fork: chain 0 continues on `st`, chains 1.. run on their own streams behind an event (captured as graph branches)
    if (G > 1) {
        RDM_CHECK_CUDA(cudaEventRecord(n->ev_fork, st));
        for (int g = 1; g < G; g++) RDM_CHECK_CUDA(cudaStreamWaitEvent(n->side[g], n->ev_fork, 0));
    }
    int rc = RDM_OK;
    for (int g = 0, b0 = 0; g < G; g++) {
        const int nb = B2 / G + (g < B2 % G ? 1 : 0);
        Arena ca; ca.dry = false; ca.base = A.base + n->shared_bytes + (size_t)g * n->chain_bytes; ca.cap = n->chain_bytes;
        Ctx cc{n, g == 0 ? st : n->side[g], false}; cc.A = &ca; cc.stats_off = (size_t)g * n->chain_stats; cc.b0 = b0;
        gemm_tc_set_workspace_slot(g);
        const int r = chain_impl(cc, x_nchw, Bx, t, nb, H, W, emb_all, out_nchw);
        if (r != RDM_OK && rc == RDM_OK) rc = r;          // keep issuing: the join below must still happen (an open capture has to be closed cleanly)
        b0 += nb;
    }
    gemm_tc_set_workspace_slot(0);
    for (int g = 1; g < G; g++) {
        RDM_CHECK_CUDA(cudaEventRecord(n->ev_join[g], n->side[g]));
        RDM_CHECK_CUDA(cudaStreamWaitEvent(st, n->ev_join[g], 0));
    }
    return rc;
}

// ---- first-stage VQ decoder -----------------------------------------------------------------------------------
// Mirrors ldm/taming `Decoder.__init__` (conv_in, mid.block_1 / attn_1 / block_2, up levels from the coarsest, norm_out, conv_out) with the
// latent-diffusion checkpoint key layout (`decoder.*`, `quantize.embedding.weight`, `post_quant_conv.*`).  Called twice like build_net.
int add_dec_res(Net* n, const std::string& p, int cin, int cout) {
    ResW r{}; r.cin = cin; r.cout = cout; r.eps = 1e-6f; r.emb_off = -1;
    r.n1 = make_norm(n, p + ".norm1", cin); r.c1 = make_conv(n, p + ".conv1", cin, cout, 3);
    r.n2 = make_norm(n, p + ".norm2", cout); r.c2 = make_conv(n, p + ".conv2", cout, cout, 3);
    r.has_skip = cin != cout;
    if (r.has_skip) { r.skip = make_conv(n, p + ".nin_shortcut", cin, cout, 1); r.b2s = walloc(n, cout); }
    n->res.push_back(r);
    return (int)n->res.size() - 1;
}
int add_dec_attn(Net* n, const std::string& p, int C) {
    AttnW a{}; a.C = C;
    a.norm = make_norm(n, p + ".norm", C);
    a.qkv.in = C; a.qkv.out = 3 * C; a.qkv.w = walloc(n, (size_t)3 * C * C); a.qkv.b = walloc(n, 3 * C);
    const char* nm[3] = {".q", ".k", ".v"};
    for (int i = 0; i < 3; i++) {
        reg_rows(n, p + nm[i] + ".weight", a.qkv.w, C, C, i * C, 1);          // 1x1 conv weight [C, C, 1, 1] == [C][C]
        reg_rows(n, p + nm[i] + ".bias", a.qkv.b, C, 1, i * C, 1);
    }
    a.proj_out = make_conv(n, p + ".proj_out", C, C, 1);
    n->attns.push_back(a);
    return (int)n->attns.size() - 1;
}
void build_decoder(Net* n) {
    const rdm_vqdec_cfg& c = n->dcfg;
    n->woff = 0; n->params.clear(); n->param_order.clear(); n->res.clear(); n->attns.clear(); n->convs.clear(); n->dec_layers.clear();
    const int L = c.n_ch_mult;
    int block_in = c.ch * c.ch_mult[L - 1], curr_res = c.resolution >> (L - 1);
    n->dec_conv_in = make_conv(n, "decoder.conv_in", c.z_channels, block_in, 3);
    n->dec_layers.push_back({D_RES, add_dec_res(n, "decoder.mid.block_1", block_in, block_in)});
    n->dec_layers.push_back({D_ATTN, add_dec_attn(n, "decoder.mid.attn_1", block_in)});
    n->dec_layers.push_back({D_RES, add_dec_res(n, "decoder.mid.block_2", block_in, block_in)});
    for (int lvl = L - 1; lvl >= 0; lvl--) {
        const int block_out = c.ch * c.ch_mult[lvl];
        const std::string up = "decoder.up." + std::to_string(lvl);
        for (int i = 0; i <= c.num_res_blocks; i++) {
            n->dec_layers.push_back({D_RES, add_dec_res(n, up + ".block." + std::to_string(i), block_in, block_out)});
            block_in = block_out;
            if (in_list(c.attn_resolutions, c.n_attn_resolutions, curr_res))
                n->dec_layers.push_back({D_ATTN, add_dec_attn(n, up + ".attn." + std::to_string(i), block_in)});
        }
        if (lvl != 0) {
            n->convs.push_back(make_up_conv(n, up + ".upsample.conv", block_in, block_in));
            n->dec_layers.push_back({D_UP, (int)n->convs.size() - 1});
            curr_res *= 2;
        }
    }
    n->dec_norm_out = make_norm(n, "decoder.norm_out", block_in);
    n->dec_conv_out = make_conv(n, "decoder.conv_out", block_in, c.out_ch, 3);
    n->codebook = walloc(n, (size_t)c.n_embed * c.embed_dim); reg_plain(n, "quantize.embedding.weight", n->codebook, (size_t)c.n_embed * c.embed_dim);
    n->pq_w = walloc(n, (size_t)c.z_channels * c.embed_dim); reg_plain(n, "post_quant_conv.weight", n->pq_w, (size_t)c.z_channels * c.embed_dim);
    n->pq_b = walloc(n, c.z_channels); reg_plain(n, "post_quant_conv.bias", n->pq_b, c.z_channels);
}

// AttnBlock.forward: h = GN(x); q,k,v = 1x1 convs; w = softmax(q k^T * C^-0.5) over ALL pixels of the image; out = x + proj_out(w v).
// One head of width C (512): both contractions are real GEMMs (4096 x 4096 x 512 per image at 64 x 64), so they run on the tcgen05 engine
// per image with the K / V^T planes of that image as the "weight" operand; softmax and the V transpose are row / tile kernels in between.
void run_dec_attn(Ctx& cx, const AttnW& a, const Act& x, View out) {
    Net* n = cx.n; Arena& A = *cx.A; size_t mk = A.mark();
    const int M = x.M(), C = a.C, HW = x.H * x.W;
    Opnd g = fresh_opnd(cx, M, C, true);
    gn(cx, x, a.norm, 1e-6f, 0, g);
    Opnd qkv; qkv.hi = (__nv_bfloat16*)A.alloc((size_t)M * 3 * C * 2); qkv.ldb = 3 * C; qkv.f.C = 3 * C; qkv.f16 = 1;
    lin_any(cx, g, M, a.qkv, GemmEpi(), qkv);
    Opnd o; o.hi = (__nv_bfloat16*)A.alloc((size_t)M * C * 2); o.ldb = C; o.f.C = C; o.f16 = 1;
    float* S = A.allocf((size_t)HW * HW);
    __half* P = (__half*)A.alloc((size_t)HW * HW * 2);
    __half* Vt = (__half*)A.alloc((size_t)C * HW * 2);
    const float scale = 1.f / sqrtf((float)C);
    for (int b = 0; b < x.B; b++) {
        const __nv_bfloat16* base = qkv.hi + (size_t)b * HW * 3 * C;
        TcA qa; qa.hi = base; qa.ld = 3 * C; qa.B = 1; qa.H = 1; qa.W = HW; qa.C = C; qa.ksize = 1;
        TcW kw; kw.hi = base + C; kw.ld = 3 * C; kw.N = HW; kw.K = C; kw.dynamic = 1;
        { GemmEpi e; e.out = S; e.out_ld = HW; RUN(gemm_tc(qa, kw, e, nullptr, nullptr, 0, 1, 1, cx.st)); }
        RUN(k_softmax_rows(S, HW, HW, HW, scale, P, HW, cx.st));
        RUN(k_transpose_plane(reinterpret_cast<const __half*>(base + 2 * C), 3 * C, HW, C, Vt, HW, cx.st));
        TcA pa; pa.hi = reinterpret_cast<const __nv_bfloat16*>(P); pa.ld = HW; pa.B = 1; pa.H = 1; pa.W = HW; pa.C = HW; pa.ksize = 1;
        TcW vw; vw.hi = reinterpret_cast<const __nv_bfloat16*>(Vt); vw.ld = HW; vw.N = C; vw.K = HW; vw.dynamic = 1;
        RUN(gemm_tc(pa, vw, GemmEpi(), o.hi + (size_t)b * HW * C, nullptr, C, 1, 1, cx.st));
    }
    { GemmEpi e; e.res = x.v.p; e.res_ld = x.v.ld; conv_any(cx, o, x, a.proj_out, e, from_view(out)); }
    A.release(mk);
}

// images NCHW [B, out_ch, 2^(L-1) h, 2^(L-1) w] = Decoder(post_quant_conv(quantize(z)))  for z NCHW [B, embed_dim, h, w]
int dec_forward_impl(Net* n, const float* z_nchw, int B, int h, int w, int quantize, float* out_nchw, cudaStream_t st, bool dry) {
    Ctx cx{n, st, dry};
    Arena& A = n->arena; A.off = 0; A.dry = dry; n->stats_off = 0;
    cx.A = &A;
    const rdm_vqdec_cfg& c = n->dcfg;
    if (!dry) RDM_CHECK_CUDA(cudaMemsetAsync(n->stats, 0, n->stats_cap * sizeof(double), st));
    // two ping-pong activation buffers sized for the largest layer output
    size_t big = 0;
    { int H = h, W = w, ch = c.ch * c.ch_mult[c.n_ch_mult - 1];
      big = (size_t)B * H * W * ch;
      for (const DecLayer& L : n->dec_layers) {
          if (L.kind == D_RES) ch = n->res[L.idx].cout;
          if (L.kind == D_UP) { H *= 2; W *= 2; }
          big = std::max(big, (size_t)B * H * W * ch);
      } }
    float* buf[2] = {A.allocf(big), A.allocf(big)};
    int H = h, W = w, cur = 0;
    Act x{View(buf[cur], n->dec_conv_in.cout, n->dec_conv_in.cout), B, H, W};
    if (c.embed_dim <= 4) {
        View z0 = fresh(cx, B * H * W, 4); z0.C = c.z_channels;
        RUN(k_vq_quantize(z_nchw, B, c.embed_dim, H * W, n->codebook, c.n_embed, n->pq_w, n->pq_b, c.z_channels, quantize, z0, st));
        gemm_any(cx, from_view(z0), B, H, W, c.z_channels, 3, 1, 0, n->dec_conv_in.w, n->dec_conv_in.b, n->dec_conv_in.cout, GemmEpi(), from_view(x.v));
    } else {
        // wide latents (taming VQGAN-f16 of the RARM models: embed_dim = z_channels = 256): z already holds the codebook entries
        // (`quantize.get_codebook_entry`, taming cond_transformer.decode_to_img), so post_quant_conv (1x1) and conv_in (3x3) are real GEMMs
        const int M = B * H * W;
        size_t mk = A.mark();
        View zf = fresh(cx, M, c.embed_dim);
        RUN(k_nchw_to_nhwc(z_nchw, B, B, c.embed_dim, H, W, zf, st));
        Opnd za = fresh_opnd(cx, M, c.embed_dim, true);
        RUN(k_split_planes(zf, M, za.out4(), st));
        Opnd zq = fresh_opnd(cx, M, c.z_channels, true);
        gemm_any(cx, za, B, H, W, c.embed_dim, 1, 1, 0, n->pq_w, n->pq_b, c.z_channels, GemmEpi(), zq);
        gemm_any(cx, zq, B, H, W, c.z_channels, 3, 1, 0, n->dec_conv_in.w, n->dec_conv_in.b, n->dec_conv_in.cout, GemmEpi(), from_view(x.v));
        A.release(mk);
    }
    for (const DecLayer& L : n->dec_layers) {
        const int nxt = cur ^ 1;
        if (L.kind == D_RES) {
            const ResW& r = n->res[L.idx];
            View o(buf[nxt], r.cout, r.cout);
            run_res(cx, r, x, nullptr, o);
            x = Act{o, B, H, W};
        } else if (L.kind == D_ATTN) {
            View o(buf[nxt], x.v.C, x.v.C);
            run_dec_attn(cx, n->attns[L.idx], x, o);
            x = Act{o, B, H, W};
        } else {
            const Conv& cv = n->convs[L.idx];
            View o(buf[nxt], cv.cout, cv.cout);
            run_up(cx, cv, x, o);
            H *= 2; W *= 2;
            x = Act{o, B, H, W};
        }
        cur = nxt;
    }
    Opnd a = fresh_opnd(cx, x.M(), x.v.C, tc_ok(cx, B, H, W, x.v.C, 3));
    gn(cx, x, n->dec_norm_out, 1e-6f, 1, a);
    View o = fresh(cx, x.M(), 4); o.C = c.out_ch;
    conv_any(cx, a, x, n->dec_conv_out, GemmEpi(), from_view(o));
    RUN(k_nhwc_to_nchw(o, B, c.out_ch, H, W, out_nchw, st));
    n->stats_off = cx.stats_off;
    return cx.rc;
}

// (re)build the bf16 hi/lo planes of every packed weight matrix: one element-wise pass over the weight arena
int ensure_weight_planes(Net* n, cudaStream_t st) {
    if (n->mode == RDM_UNET_MODE_FP32 || !n->planes_dirty) return RDM_OK;
    if (!n->wb_hi) RDM_CHECK_CUDA(cudaMalloc((void**)&n->wb_hi, n->wfloats * 2));
    if (!n->wb_lo) RDM_CHECK_CUDA(cudaMalloc((void**)&n->wb_lo, n->wfloats * 2));
    for (const ResW& r : n->res) if (r.b2s) RDM_TRY(k_add_vec(r.c2.b, r.skip.b, r.b2s, r.cout, st));       // bias of the K-extended second conv (run_res)
    for (const Conv& c : n->convs) if (c.wfold) RDM_TRY(k_fold_up_weights(c.w, c.cout, c.cin, c.wfold, st));    // Upsample + conv as four 2x2-tap convs (run_up)
    RDM_TRY(k_split_planes(View(n->wbase, 64, 64), (long long)(n->wfloats / 64), Out4(n->wb_hi, n->wb_lo, 64, mode_f16(n->mode)), st));
    n->planes_dirty = false;
    return RDM_OK;
}

int ensure_plan(Net* n, int B2, int H, int W) {
    const int G = effective_chains(n, B2);
    if (n->plan_B == B2 && n->plan_H == H && n->plan_W == W && n->plan_chains == G && n->arena.base) return RDM_OK;
    if (G > 1 && !n->ev_fork) {
        RDM_CHECK_CUDA(cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming));
        for (int g = 1; g < 8; g++) {
            RDM_CHECK_CUDA(cudaStreamCreateWithFlags(&n->side[g], cudaStreamNonBlocking));
            RDM_CHECK_CUDA(cudaEventCreateWithFlags(&n->ev_join[g], cudaEventDisableTiming));
        }
    }
    n->plan_chains = G;
    n->arena.dry = true; n->arena.off = 0; n->arena.peak = 0; n->stats_off = 0;
    RDM_TRY(forward_impl(n, nullptr, B2, nullptr, B2, H, W, nullptr, 0, true));
    size_t need = n->arena.peak, sneed = n->stats_off;
    if (need > n->arena.cap) {
        if (n->arena.base) cudaFree(n->arena.base);
        n->arena.base = nullptr; n->arena.cap = 0;
        RDM_CHECK_CUDA(cudaMalloc((void**)&n->arena.base, need));
        n->arena.cap = need;
    }
    if (sneed > n->stats_cap) {
        if (n->stats) cudaFree(n->stats);
        n->stats = nullptr; n->stats_cap = 0;
        RDM_CHECK_CUDA(cudaMalloc((void**)&n->stats, sneed * sizeof(double)));
        n->stats_cap = sneed;
    }
    n->plan_B = B2; n->plan_H = H; n->plan_W = W;
    if (n->fwd_exec) { cudaGraphExecDestroy(n->fwd_exec); n->fwd_exec = nullptr; }      // buffers moved: recapture
    if (n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }
    return RDM_OK;
}


int ensure_io(Net* n, int B2, int H, int W) {
    const size_t C = (size_t)(n->cfg.in_channels > n->cfg.out_channels ? n->cfg.in_channels : n->cfg.out_channels);
    const size_t need = (size_t)B2 * C * H * W;
    if (!n->cap_stream) RDM_CHECK_CUDA(cudaStreamCreateWithFlags(&n->cap_stream, cudaStreamNonBlocking));
    if (!n->step_dev) RDM_CHECK_CUDA(cudaMalloc((void**)&n->step_dev, sizeof(int)));
    if ((size_t)B2 > n->tin_cap) {          // the timestep vector has its own capacity: a larger batch of SMALLER images must not reuse a short one
        if (n->t_in) cudaFree(n->t_in);
        n->t_in = nullptr; n->tin_cap = 0;
        if (n->fwd_exec) { cudaGraphExecDestroy(n->fwd_exec); n->fwd_exec = nullptr; }       // the captured graphs hold the old pointer
        if (n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }
        RDM_CHECK_CUDA(cudaMalloc((void**)&n->t_in, (size_t)B2 * 8));
        n->tin_cap = (size_t)B2;
    }
    if (need <= n->io_cap) return RDM_OK;
    for (float** p : {&n->x_in, &n->eps_buf, &n->x_state, &n->p0_buf}) { if (*p) cudaFree(*p); *p = nullptr; }
    n->io_cap = 0;
    if (n->fwd_exec) { cudaGraphExecDestroy(n->fwd_exec); n->fwd_exec = nullptr; }
    if (n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->x_in, need * 4)); RDM_CHECK_CUDA(cudaMalloc((void**)&n->eps_buf, need * 4));
    RDM_CHECK_CUDA(cudaMalloc((void**)&n->x_state, need * 4)); RDM_CHECK_CUDA(cudaMalloc((void**)&n->p0_buf, need * 4));
    n->io_cap = need;
    return RDM_OK;
}

// Captures `body` (which must only enqueue work on n->cap_stream) into an executable graph.
template <typename F> int capture_graph(Net* n, cudaGraphExec_t* exec, unsigned long long* nkernels, F body) {
    if (*exec) { cudaGraphExecDestroy(*exec); *exec = nullptr; }
    RDM_CHECK_CUDA(cudaStreamBeginCapture(n->cap_stream, cudaStreamCaptureModeThreadLocal));
    const unsigned long long before = g_rdm_launches;
    int rc = body(n->cap_stream);
    *nkernels = g_rdm_launches - before; g_rdm_launches = before;       // captured, not executed
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamEndCapture(n->cap_stream, &g);
    if (rc != RDM_OK) { if (g) cudaGraphDestroy(g); return rc; }
    RDM_REQUIRE(ce == cudaSuccess && g, RDM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    RDM_REQUIRE(ce == cudaSuccess, RDM_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    return RDM_OK;
}

// eps_buf = UNet(x_in [Bx], t_in [B2]) by graph replay on `st` (first call per shape runs eagerly once, then captures).
int forward_staged(Net* n, int Bx, int B2, int H, int W, cudaStream_t st) {
    if (!n->use_graph || n->debug || n->profile) return forward_impl(n, n->x_in, Bx, n->t_in, B2, H, W, n->eps_buf, st, false);
    const int key[6] = {Bx, B2, H, W, n->mode, n->skip};
    if (!n->fwd_exec || memcmp(key, n->fwd_key, sizeof(key)) != 0) {
        RDM_TRY(forward_impl(n, n->x_in, Bx, n->t_in, B2, H, W, n->eps_buf, st, false));     // eager warm-up: sets kernel attributes
        RDM_CHECK_CUDA(cudaStreamSynchronize(st));
        RDM_TRY(capture_graph(n, &n->fwd_exec, &n->fwd_kernels, [&](cudaStream_t cs) { return forward_impl(n, n->x_in, Bx, n->t_in, B2, H, W, n->eps_buf, cs, false); }));
        memcpy(n->fwd_key, key, sizeof(key));
        return RDM_OK;                         // the warm-up already produced this call's result
    }
    RDM_CHECK_CUDA(cudaGraphLaunch(n->fwd_exec, st));
    g_rdm_launches += n->fwd_kernels;
    return RDM_OK;
}

}  // namespace

extern "C" {

int rdm_unet_create(rdm_unet_t** out, const rdm_unet_cfg* cfg, int32_t device) {
    RDM_REQUIRE(out && cfg, RDM_ERR_ARG, "rdm_unet_create: null argument");
    RDM_REQUIRE(cfg->n_channel_mult >= 1 && cfg->n_channel_mult <= 8 && cfg->n_attention_resolutions >= 0 && cfg->n_attention_resolutions <= 8,
                RDM_ERR_ARG, "rdm_unet_create: bad list sizes");
    RDM_REQUIRE(cfg->model_channels % 32 == 0, RDM_ERR_UNSUPPORTED, "rdm_unet_create: model_channels must be a multiple of 32 (GroupNorm32)");
    RDM_REQUIRE(cfg->num_head_channels == 32, RDM_ERR_UNSUPPORTED, "rdm_unet_create: only num_head_channels=32 is implemented (got %d)", cfg->num_head_channels);
    RDM_REQUIRE(cfg->transformer_depth == 1, RDM_ERR_UNSUPPORTED, "rdm_unet_create: transformer_depth must be 1");
    RDM_REQUIRE(cfg->context_dim > 0 && cfg->context_dim % 16 == 0, RDM_ERR_UNSUPPORTED, "rdm_unet_create: context_dim %d", cfg->context_dim);
    DeviceGuard guard(device);
    RDM_REQUIRE(guard.ok, RDM_ERR_CUDA, "rdm_unet_create: cannot select device %d", device);
    rdm_unet* n = new rdm_unet();
    n->device = device; n->cfg = *cfg;
    build_net(n);                                   // counting pass
    n->wfloats = n->woff;
    if (cudaMalloc((void**)&n->wbase, n->wfloats * sizeof(float)) != cudaSuccess) { delete n; rdm_set_error("rdm_unet_create: cudaMalloc(%zu) for weights failed", n->wfloats * 4); return RDM_ERR_CUDA; }
    cudaMemset(n->wbase, 0, n->wfloats * sizeof(float));
    build_net(n);                                   // placing pass
    // context K/V offsets
    size_t off = 0; n->ctx_off.resize(n->sts.size());
    for (auto& s : n->sts) { n->ctx_off[s.id] = off; off += (size_t)2 * s.C; }      // per context row; scaled by rows in set_context
    *out = n;
    return RDM_OK;
}

void rdm_unet_destroy(rdm_unet_t* n) {
    if (!n) return;
    DeviceGuard guard(n->device);
    if (n->wbase) cudaFree(n->wbase);
    if (n->wb_hi) cudaFree(n->wb_hi);
    if (n->wb_lo) cudaFree(n->wb_lo);
    if (n->arena.base) cudaFree(n->arena.base);
    if (n->stats) cudaFree(n->stats);
    if (n->ctx_kv) cudaFree(n->ctx_kv);
    if (n->t_dev) cudaFree(n->t_dev);
    if (n->fwd_exec) cudaGraphExecDestroy(n->fwd_exec);
    if (n->step_exec) cudaGraphExecDestroy(n->step_exec);
    for (float* p : {n->x_in, n->eps_buf, n->x_state, n->p0_buf}) if (p) cudaFree(p);
    if (n->t_in) cudaFree(n->t_in);
    if (n->step_dev) cudaFree(n->step_dev);
    if (n->cap_stream) cudaStreamDestroy(n->cap_stream);
    if (n->ev_fork) cudaEventDestroy(n->ev_fork);
    for (int g = 1; g < 8; g++) { if (n->side[g]) cudaStreamDestroy(n->side[g]); if (n->ev_join[g]) cudaEventDestroy(n->ev_join[g]); }
    if (n->dec_exec) cudaGraphExecDestroy(n->dec_exec);
    if (n->dec_in) cudaFree(n->dec_in);
    if (n->dec_out) cudaFree(n->dec_out);
    delete n;
}

int64_t rdm_unet_num_params(const rdm_unet_t* n) { return n ? (int64_t)n->param_order.size() : 0; }
const char* rdm_unet_param_name(const rdm_unet_t* n, int64_t i) { return (n && i >= 0 && i < (int64_t)n->param_order.size()) ? n->param_order[i].c_str() : nullptr; }
int64_t rdm_unet_param_numel(const rdm_unet_t* n, const char* name) {
    if (!n || !name) return -1;
    auto it = n->params.find(name);
    return it == n->params.end() ? -1 : (int64_t)it->second.numel;
}

int rdm_unet_load(rdm_unet_t* n, const char* name, const float* host, int64_t numel) {
    RDM_REQUIRE(n && name && host, RDM_ERR_ARG, "rdm_unet_load: null argument");
    auto it = n->params.find(name);
    RDM_REQUIRE(it != n->params.end(), RDM_ERR_ARG, "rdm_unet_load: unknown parameter '%s'", name);
    ParamSlot& s = it->second;
    RDM_REQUIRE((size_t)numel == s.numel, RDM_ERR_ARG, "rdm_unet_load: '%s' has %lld elements, expected %zu", name, (long long)numel, s.numel);
    DeviceGuard guard(n->device);
    if (s.rows > 0) {
        RDM_CHECK_CUDA(cudaMemcpy2D(s.dst + (size_t)s.dst_row0 * s.row_len, (size_t)s.dst_row_step * s.row_len * sizeof(float), host,
                                    (size_t)s.row_len * sizeof(float), (size_t)s.row_len * sizeof(float), s.rows, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> packed;
        s.pack(host, packed);
        const float* src = packed.empty() ? host : packed.data();
        RDM_CHECK_CUDA(cudaMemcpy(s.dst, src, s.dst_numel * sizeof(float), cudaMemcpyHostToDevice));
    }
    s.loaded = true;
    n->ctx_B = 0;           // projections of a previously set context are stale
    n->planes_dirty = true;
    return RDM_OK;
}

int64_t rdm_unet_missing(const rdm_unet_t* n) {
    if (!n) return -1;
    int64_t m = 0; for (auto& kv : n->params) if (!kv.second.loaded) m++;
    return m;
}

int rdm_unet_set_debug(rdm_unet_t* n, int32_t on) { RDM_REQUIRE(n, RDM_ERR_ARG, "rdm_unet_set_debug: null handle"); n->debug = on; return RDM_OK; }
const char* rdm_unet_debug_log(const rdm_unet_t* n) { return n ? n->debug_log.c_str() : ""; }

int rdm_unet_set_mode(rdm_unet_t* n, int32_t mode) {
    RDM_REQUIRE(n, RDM_ERR_ARG, "rdm_unet_set_mode: null handle");
    RDM_REQUIRE(mode >= RDM_UNET_MODE_FP32 && mode <= RDM_UNET_MODE_TC_FP16, RDM_ERR_UNSUPPORTED, "rdm_unet_set_mode: unknown mode %d", mode);
    if (mode != n->mode) { n->plan_B = 0; n->planes_dirty = true; }      // workspace layout depends on the engine
    n->mode = mode; return RDM_OK;
}

int rdm_unet_set_context(rdm_unet_t* n, const float* ctx, int32_t B2, int32_t k, void* stream) {
    RDM_REQUIRE(n && ctx, RDM_ERR_ARG, "rdm_unet_set_context: null argument");
    RDM_REQUIRE(B2 >= 1 && k >= 1, RDM_ERR_ARG, "rdm_unet_set_context: B2=%d k=%d", B2, k);
    RDM_REQUIRE(rdm_unet_missing(n) == 0, RDM_ERR_STATE, "rdm_unet_set_context: %lld parameters not loaded", (long long)rdm_unet_missing(n));
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = B2 * k;
    size_t per_row = 0; for (auto& s : n->sts) per_row += (size_t)2 * s.C;
    size_t need = per_row * rows;
    bool relayout = (B2 != n->ctx_B || k != n->ctx_k);
    if (need > n->ctx_kv_floats) {
        if (n->ctx_kv) cudaFree(n->ctx_kv);
        n->ctx_kv = nullptr; n->ctx_kv_floats = 0;
        RDM_CHECK_CUDA(cudaMalloc((void**)&n->ctx_kv, need * sizeof(float)));
        n->ctx_kv_floats = need; relayout = true;
    }
    if (relayout) {       // captured graphs hold the K/V addresses and the neighbour count
        if (n->fwd_exec) { cudaGraphExecDestroy(n->fwd_exec); n->fwd_exec = nullptr; }
        if (n->step_exec) { cudaGraphExecDestroy(n->step_exec); n->step_exec = nullptr; }
    }
    size_t off = 0;
    Ctx cx{n, st, false}; cx.A = &n->arena;
    for (auto& s : n->sts) {
        n->ctx_off[s.id] = off;
        View dst(n->ctx_kv + off, 2 * s.C, 2 * s.C);
        gemm_any(cx, from_view(View(const_cast<float*>(ctx), n->cfg.context_dim, n->cfg.context_dim)), rows, 1, 1, s.kv2.in, 1, 1, 0,
                 s.kv2.w, nullptr, s.kv2.out, GemmEpi(), from_view(dst));
        off += (size_t)2 * s.C * rows;
    }
    n->ctx_B = B2; n->ctx_k = k;
    return cx.rc;
}

int rdm_unet_forward(rdm_unet_t* n, const float* x, int32_t Bx, const int64_t* t, int32_t B2, int32_t H, int32_t W, float* eps_out, void* stream) {
    RDM_REQUIRE(n && x && t && eps_out, RDM_ERR_ARG, "rdm_unet_forward: null argument");
    RDM_REQUIRE(n->kind == 0, RDM_ERR_ARG, "rdm_unet_forward: this handle is a VQ decoder (use rdm_vqdec_decode)");
    RDM_REQUIRE(B2 >= 1 && (Bx == B2 || Bx * 2 == B2), RDM_ERR_ARG, "rdm_unet_forward: Bx=%d must equal B2=%d or B2/2", Bx, B2);
    RDM_REQUIRE(n->ctx_B == B2, RDM_ERR_STATE, "rdm_unet_forward: context was set for batch %d, forward called with %d (call rdm_unet_set_context first)", n->ctx_B, B2);
    int div = 1; for (int i = 1; i < n->cfg.n_channel_mult; i++) div *= 2;
    RDM_REQUIRE(H % div == 0 && W % div == 0, RDM_ERR_ARG, "rdm_unet_forward: H=%d W=%d must be multiples of %d", H, W, div);
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    RDM_TRY(ensure_plan(n, B2, H, W));
    RDM_TRY(ensure_weight_planes(n, st));
    RDM_TRY(ensure_io(n, B2, H, W));
    const size_t per = (size_t)H * W;
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->x_in, x, (size_t)Bx * n->cfg.in_channels * per * 4, cudaMemcpyDeviceToDevice, st));
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->t_in, t, (size_t)B2 * 8, cudaMemcpyDeviceToDevice, st));
    RDM_TRY(forward_staged(n, Bx, B2, H, W, st));
    RDM_CHECK_CUDA(cudaMemcpyAsync(eps_out, n->eps_buf, (size_t)B2 * n->cfg.out_channels * per * 4, cudaMemcpyDeviceToDevice, st));
    return RDM_OK;
}

int rdm_unet_set_ablation(rdm_unet_t* n, int32_t mask) { RDM_REQUIRE(n, RDM_ERR_ARG, "rdm_unet_set_ablation: null handle"); n->skip = mask; return RDM_OK; }
int rdm_unet_set_chains(rdm_unet_t* n, int32_t chains) {
    RDM_REQUIRE(n, RDM_ERR_ARG, "rdm_unet_set_chains: null handle");
    RDM_REQUIRE(chains >= 1 && chains <= 8, RDM_ERR_ARG, "rdm_unet_set_chains: %d outside 1..8", chains);
    n->chains = chains;                 // takes effect at the next forward (ensure_plan re-plans and drops the captured graphs)
    return RDM_OK;
}
int rdm_unet_set_graph(rdm_unet_t* n, int32_t on) { RDM_REQUIRE(n, RDM_ERR_ARG, "rdm_unet_set_graph: null handle"); n->use_graph = on; return RDM_OK; }

int rdm_unet_profile_forward(rdm_unet_t* n, const float* x, int32_t Bx, const int64_t* t, int32_t B2, int32_t H, int32_t W, float* eps_out,
                             double* out8, void* stream) {
    RDM_REQUIRE(n && out8, RDM_ERR_ARG, "rdm_unet_profile_forward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    // 1. a plain forward: validates arguments, stages the inputs, sets kernel attributes, produces eps_out
    RDM_TRY(rdm_unet_forward(n, x, Bx, t, B2, H, W, eps_out, stream));
    RDM_CHECK_CUDA(cudaStreamSynchronize(st));
    DeviceGuard guard(n->device);
    // 2. the same forward captured into a graph with external event-record nodes around every GEMM: durations are those of the
    //    graph-replayed kernels (no host launch gaps inside the brackets).  Replayed twice; the second replay is the one measured.
    n->profile = 2; n->prof_ev.clear();
    RDM_TRY(ensure_plan(n, B2, H, W));                 // per-GEMM brackets follow one stream: re-plan with a single chain (the next ordinary forward re-plans back) n->prof_flops.clear(); n->prof_kind.clear(); n->prof_desc.clear(); n->prof_text.clear();
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaGraphExec_t pexec = nullptr; unsigned long long nk = 0;
    int rc = capture_graph(n, &pexec, &nk, [&](cudaStream_t cs) {
        cudaEventRecordWithFlags(t0, cs, cudaEventRecordExternal);
        int r = forward_impl(n, n->x_in, Bx, n->t_in, B2, H, W, n->eps_buf, cs, false);
        cudaEventRecordWithFlags(t1, cs, cudaEventRecordExternal);
        return r;
    });
    n->profile = 0;
    if (rc == RDM_OK) {
        for (int rep = 0; rep < 2 && rc == RDM_OK; rep++) {
            if (cudaGraphLaunch(pexec, st) != cudaSuccess) { rdm_set_error("rdm_unet_profile_forward: graph launch failed"); rc = RDM_ERR_CUDA; }
            g_rdm_launches += nk;
        }
    }
    cudaError_t ce = cudaStreamSynchronize(st);
    for (int i = 0; i < 8; i++) out8[i] = 0.0;
    if (rc == RDM_OK && ce == cudaSuccess) {
        float ms = 0.f; cudaEventElapsedTime(&ms, t0, t1); out8[4] = ms;
        for (size_t i = 0; i < n->prof_flops.size(); i++) {
            float m = 0.f; cudaEventElapsedTime(&m, n->prof_ev[2 * i], n->prof_ev[2 * i + 1]);
            const int kind = n->prof_kind[i];
            out8[kind ? 0 : 2] += m; out8[kind ? 1 : 3] += n->prof_flops[i]; out8[kind ? 5 : 6] += 1.0;
            char line[256]; snprintf(line, sizeof(line), "%s ms=%.4f tflops=%.1f\n", n->prof_desc[i].c_str(), m, n->prof_flops[i] / (m * 1e-3) / 1e12);
            n->prof_text += line;
        }
    }
    for (cudaEvent_t e : n->prof_ev) cudaEventDestroy(e);
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    if (pexec) cudaGraphExecDestroy(pexec);
    n->prof_ev.clear(); n->prof_flops.clear(); n->prof_kind.clear();
    RDM_REQUIRE(ce == cudaSuccess, RDM_ERR_CUDA, "rdm_unet_profile_forward: %s", cudaGetErrorString(ce));
    return rc;
}


// ---- first-stage VQ decoder entry points ------------------------------------------------------------------------------------------------
int rdm_vqdec_create(rdm_unet_t** out, const rdm_vqdec_cfg* cfg, int32_t device) {
    RDM_REQUIRE(out && cfg, RDM_ERR_ARG, "rdm_vqdec_create: null argument");
    RDM_REQUIRE(cfg->n_ch_mult >= 1 && cfg->n_ch_mult <= 8 && cfg->n_attn_resolutions >= 0 && cfg->n_attn_resolutions <= 8, RDM_ERR_ARG, "rdm_vqdec_create: bad list sizes");
    RDM_REQUIRE(cfg->ch % 64 == 0, RDM_ERR_UNSUPPORTED, "rdm_vqdec_create: ch must be a multiple of 64 (GroupNorm32 + 64-wide K blocks), got %d", cfg->ch);
    const bool narrow = cfg->embed_dim >= 1 && cfg->embed_dim <= 4 && cfg->z_channels >= 1 && cfg->z_channels <= 4;
    const bool wide = cfg->embed_dim >= 64 && cfg->embed_dim % 64 == 0 && cfg->z_channels >= 64 && cfg->z_channels % 64 == 0;      // taming VQGAN (RARM): decode without quantisation only
    RDM_REQUIRE((narrow || wide) && cfg->out_ch >= 1 && cfg->out_ch <= 4, RDM_ERR_UNSUPPORTED,
                "rdm_vqdec_create: embed_dim / z_channels must both be 1..4 or both multiples of 64, out_ch 1..4 (got %d / %d / %d)", cfg->embed_dim, cfg->z_channels, cfg->out_ch);
    RDM_REQUIRE(cfg->n_embed >= 1 && cfg->num_res_blocks >= 1, RDM_ERR_ARG, "rdm_vqdec_create: n_embed / num_res_blocks");
    DeviceGuard guard(device);
    RDM_REQUIRE(guard.ok, RDM_ERR_CUDA, "rdm_vqdec_create: cannot select device %d", device);
    rdm_unet* n = new rdm_unet();
    n->device = device; n->kind = 1; n->dcfg = *cfg; n->mode = RDM_UNET_MODE_TC_FP16X2;
    build_decoder(n);                               // counting pass
    n->wfloats = n->woff;
    if (cudaMalloc((void**)&n->wbase, n->wfloats * sizeof(float)) != cudaSuccess) { delete n; rdm_set_error("rdm_vqdec_create: cudaMalloc for weights failed"); return RDM_ERR_CUDA; }
    cudaMemset(n->wbase, 0, n->wfloats * sizeof(float));
    build_decoder(n);                               // placing pass
    *out = n;
    return RDM_OK;
}

int rdm_vqdec_decode(rdm_unet_t* n, const float* z, int32_t B, int32_t h, int32_t w, int32_t quantize, float* out, void* stream) {
    RDM_REQUIRE(n && z && out, RDM_ERR_ARG, "rdm_vqdec_decode: null argument");
    RDM_REQUIRE(n->kind == 1, RDM_ERR_ARG, "rdm_vqdec_decode: this handle is not a VQ decoder");
    RDM_REQUIRE(mode_f16(n->mode), RDM_ERR_UNSUPPORTED, "rdm_vqdec_decode: the decoder runs in the fp16 tensor-core modes only (fp16x2 / fp16), mode is %d", n->mode);
    RDM_REQUIRE(B >= 1 && h >= 1 && w >= 1, RDM_ERR_ARG, "rdm_vqdec_decode: B=%d h=%d w=%d", B, h, w);
    RDM_REQUIRE(n->dcfg.embed_dim <= 4 || quantize == 0, RDM_ERR_UNSUPPORTED, "rdm_vqdec_decode: wide codebooks (embed_dim %d) are decoded from codebook entries only (quantize = 0)", n->dcfg.embed_dim);
    for (auto& kv : n->params) RDM_REQUIRE(kv.second.loaded, RDM_ERR_STATE, "rdm_vqdec_decode: parameter '%s' was never loaded", kv.first.c_str());
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int up = 1 << (n->dcfg.n_ch_mult - 1);
    const size_t zin = (size_t)B * n->dcfg.embed_dim * h * w, zout = (size_t)B * n->dcfg.out_ch * h * w * up * up;
    if (n->plan_B != B || n->plan_H != h || n->plan_W != w || !n->arena.base) {
        n->arena.dry = true; n->arena.off = 0; n->arena.peak = 0; n->stats_off = 0;
        RDM_TRY(dec_forward_impl(n, nullptr, B, h, w, quantize, nullptr, 0, true));
        const size_t need = n->arena.peak, sneed = n->stats_off;
        if (need > n->arena.cap) {
            if (n->arena.base) cudaFree(n->arena.base);
            n->arena.base = nullptr; n->arena.cap = 0;
            RDM_CHECK_CUDA(cudaMalloc((void**)&n->arena.base, need));
            n->arena.cap = need;
        }
        if (sneed > n->stats_cap) {
            if (n->stats) cudaFree(n->stats);
            n->stats = nullptr; n->stats_cap = 0;
            RDM_CHECK_CUDA(cudaMalloc((void**)&n->stats, sneed * sizeof(double)));
            n->stats_cap = sneed;
        }
        n->plan_B = B; n->plan_H = h; n->plan_W = w;
        if (n->dec_exec) { cudaGraphExecDestroy(n->dec_exec); n->dec_exec = nullptr; }
    }
    RDM_TRY(ensure_weight_planes(n, st));
    if (!n->cap_stream) RDM_CHECK_CUDA(cudaStreamCreateWithFlags(&n->cap_stream, cudaStreamNonBlocking));
    if (zin + zout > n->dec_io_cap) {
        if (n->dec_in) cudaFree(n->dec_in);
        if (n->dec_out) cudaFree(n->dec_out);
        n->dec_in = n->dec_out = nullptr; n->dec_io_cap = 0;
        if (n->dec_exec) { cudaGraphExecDestroy(n->dec_exec); n->dec_exec = nullptr; }
        RDM_CHECK_CUDA(cudaMalloc((void**)&n->dec_in, zin * 4)); RDM_CHECK_CUDA(cudaMalloc((void**)&n->dec_out, zout * 4));
        n->dec_io_cap = zin + zout;
    }
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->dec_in, z, zin * 4, cudaMemcpyDeviceToDevice, st));
    const int key[5] = {B, h, w, quantize, n->mode};
    if (!n->use_graph) {
        RDM_TRY(dec_forward_impl(n, n->dec_in, B, h, w, quantize, n->dec_out, st, false));
    } else if (!n->dec_exec || memcmp(key, n->dec_key, sizeof(key)) != 0) {
        RDM_TRY(dec_forward_impl(n, n->dec_in, B, h, w, quantize, n->dec_out, st, false));       // eager warm-up: sets kernel attributes, produces this call's result
        RDM_CHECK_CUDA(cudaStreamSynchronize(st));
        RDM_TRY(capture_graph(n, &n->dec_exec, &n->dec_kernels, [&](cudaStream_t cs) { return dec_forward_impl(n, n->dec_in, B, h, w, quantize, n->dec_out, cs, false); }));
        memcpy(n->dec_key, key, sizeof(key));
    } else {
        RDM_CHECK_CUDA(cudaGraphLaunch(n->dec_exec, st));
        g_rdm_launches += n->dec_kernels;
    }
    RDM_CHECK_CUDA(cudaMemcpyAsync(out, n->dec_out, zout * 4, cudaMemcpyDeviceToDevice, st));
    return RDM_OK;
}

const char* rdm_unet_profile_text(const rdm_unet_t* n) { return n ? n->prof_text.c_str() : ""; }

int rdm_ddim_sample(rdm_unet_t* n, float* x_dev, int32_t B, int32_t H, int32_t W, const int64_t* timesteps_dev, const float* coef_dev,
                    int32_t first_step, int32_t num_steps, float cfg_scale, const float* noise_dev, float* pred_x0_dev, void* stream) {
    RDM_REQUIRE(n && x_dev && timesteps_dev && coef_dev, RDM_ERR_ARG, "rdm_ddim_sample: null argument");
    RDM_REQUIRE(B >= 1 && num_steps >= 0 && first_step >= 0, RDM_ERR_ARG, "rdm_ddim_sample: bad sizes");
    RDM_REQUIRE(n->cfg.in_channels == n->cfg.out_channels, RDM_ERR_UNSUPPORTED, "rdm_ddim_sample: eps-model needs in_channels == out_channels");
    const int cfg = cfg_scale > 1.f ? 1 : 0, B2 = cfg ? 2 * B : B;
    RDM_REQUIRE(n->ctx_B == B2, RDM_ERR_STATE, "rdm_ddim_sample: context was set for batch %d, need %d ([cond | uncond] when cfg_scale > 1)", n->ctx_B, B2);
    if (num_steps == 0) return RDM_OK;
    DeviceGuard guard(n->device);
    cudaStream_t st = (cudaStream_t)stream;
    RDM_TRY(ensure_plan(n, B2, H, W));
    RDM_TRY(ensure_weight_planes(n, st));
    RDM_TRY(ensure_io(n, B2, H, W));
    const long long nper = (long long)B * n->cfg.in_channels * H * W;
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->x_in, x_dev, (size_t)nper * 4, cudaMemcpyDeviceToDevice, st));
    RDM_CHECK_CUDA(cudaMemcpyAsync(n->step_dev, &first_step, sizeof(int), cudaMemcpyHostToDevice, st));
    // one step = {t_in <- timesteps[step]; eps = UNet(x_in); x_in <- ddim(x_in, eps, coef[step]); step++}; x_in doubles as the state
    auto body = [&](cudaStream_t cs) -> int {
        RDM_TRY(k_fill_timesteps((const long long*)timesteps_dev, n->step_dev, B2, n->t_in, cs));
        RDM_TRY(forward_impl(n, n->x_in, B, n->t_in, B2, H, W, n->eps_buf, cs, false));
        RDM_TRY(k_ddim_update_table(n->x_in, n->eps_buf, nper, cfg, cfg_scale, coef_dev, n->step_dev, noise_dev, n->x_in, n->p0_buf, cs));
        RDM_TRY(k_step_advance(n->step_dev, cs));
        return RDM_OK;
    };
    int done = 0;
    if (n->use_graph && !n->debug) {
        long long key[10] = {B, B2, H, W, n->mode, (long long)(uintptr_t)timesteps_dev, (long long)(uintptr_t)coef_dev, (long long)(uintptr_t)noise_dev, cfg,
                             (long long)(cfg_scale * 1e6)};
        if (!n->step_exec || memcmp(key, n->step_key, sizeof(key)) != 0) {
            RDM_TRY(body(st));                                 // eager first step (also the warm-up)
            RDM_CHECK_CUDA(cudaStreamSynchronize(st));
            done = 1;
            RDM_TRY(capture_graph(n, &n->step_exec, &n->step_kernels, body));
            memcpy(n->step_key, key, sizeof(key));
        }
        for (int i = done; i < num_steps; i++) { RDM_CHECK_CUDA(cudaGraphLaunch(n->step_exec, st)); g_rdm_launches += n->step_kernels; }
    } else {
        for (int i = 0; i < num_steps; i++) RDM_TRY(body(st));
    }
    RDM_CHECK_CUDA(cudaMemcpyAsync(x_dev, n->x_in, (size_t)nper * 4, cudaMemcpyDeviceToDevice, st));
    if (pred_x0_dev) RDM_CHECK_CUDA(cudaMemcpyAsync(pred_x0_dev, n->p0_buf, (size_t)nper * 4, cudaMemcpyDeviceToDevice, st));
    return RDM_OK;
}

int rdm_ddim_step(const float* x, const float* eps, int64_t n_per_half, int32_t cfg, float scale, const float* coef_dev,
                  const float* noise, float* x_prev, float* pred_x0, int32_t device, void* stream) {
    RDM_REQUIRE(x && eps && coef_dev && x_prev, RDM_ERR_ARG, "rdm_ddim_step: null argument");
    DeviceGuard guard(device);
    return k_ddim_update(x, eps, n_per_half, cfg, scale, coef_dev, noise, x_prev, pred_x0, (cudaStream_t)stream);
}

}  // extern "C"
