// SA (RT-DETR) keypoint predictor: weights, workspace and forward schedule.  Included by model.cu inside namespace spe
// (it uses that file's weight-upload helpers and the Fwd launcher wrapper); selected by spe_config::backbone == 2.
//
// Reference (second code drop, "SA"): RTDETR.forward SA/src/zoo/rtdetr/rtdetr.py:36-52 =
//   PResNet-50-vd          SA/nn/backbone/presnet.py:153-265      3x3 stem x3, max-pool, 3+4+6+3 bottlenecks (stride on
//                                                                 the 3x3, AvgPool + 1x1 shortcut), outputs at /8 /16 /32
//   HybridEncoder          SA/src/zoo/rtdetr/hybrid_encoder.py:202-401   1x1 + BN projections, one post-norm GELU
//                                                                 transformer layer on the /32 level (2-D sin-cos
//                                                                 positions), CSPRep top-down (nearest x2) and
//                                                                 bottom-up (bicubic x0.5) fusion
//   RTDETRTransformer      SA/src/zoo/rtdetr/rtdetr_decoder.py:375-777   1x1 + BN projections -> memory [B, Lv, 256],
//                                                                 enc_output + score / keypoint heads on every anchor,
//                                                                 top-k query selection, decoder layers (self-attention,
//                                                                 multi-scale deformable cross-attention, ReLU FFN) with
//                                                                 iterative keypoint refinement and a log-sigma head
// Eval mode, no denoising queries.  fp32 storage only.  Everything up to the encoder outputs runs plain TF32 (like the
// RV trunk); everything that reads the memory or the decoder state runs 3xTF32 (fp32-grade), because the top-k
// selection over 1344 anchors and the refinement chain amplify rounding noise.
//
// Same-function rewrites done at weight load (exact in real arithmetic):
//   * RepVggBlock's 3x3+BN and 1x1+BN branches are one 3x3 convolution + bias (the reference's own
//     get_equivalent_kernel_bias, hybrid_encoder.py:66-93);
//   * conv1_1 / conv1_2 have 32 output channels: their weight matrices carry 32 zero rows (the GEMM's narrowest tile
//     is 64 wide) and the next convolution reads channels [0, 32) of the 64-channel rows (GemmDesc::c_ld);
//   * the value projections of all decoder layers read the same memory: one GEMM with N = layers x 256;
//   * sampling_offsets | attention_weights of a layer share their input: one GEMM with N = 192 + 96 (+ 32 zero rows);
//   * dec_bbox_head[i].layers.0 | sigma_embed[i].layers.0 share their input: one GEMM with N = 512.

struct SaBlock {
  GemmW a, b, c, sc;       // BottleNeck: 1x1, 3x3 (stride), 1x1; BasicBlock (depth 18 / 34): 3x3 (stride), 3x3, c unused
  int cin = 0, planes = 0, stride = 1;
  bool has_short = false;
  bool basic = false;
};
struct SaCsp { GemmW c1, c2, rep, c3; };
struct SaDecLayer {
  GemmW sa_qk, sa_v, sa_out, offaw, ca_out, ff1, ff2, hb0, bb1, sg1;
  float *n1g = nullptr, *n1b = nullptr, *n2g = nullptr, *n2b = nullptr, *n3g = nullptr, *n3b = nullptr;
  float *cls_w = nullptr, *cls_b = nullptr, *bb2_w = nullptr, *bb2_b = nullptr, *sg2_w = nullptr, *sg2_b = nullptr;
};
struct SaModel {
  int R = 0, hl[3] = {0, 0, 0}, start[3] = {0, 0, 0}, Lv = 0;
  // 3xTF32 (fp32-grade products) for backbone + encoder too.  With plain TF32 there the encoder memory carries ~1e-3 of
  // relative rounding noise, which the keypoint logits (anchor logit + MLP(memory), refined through
  // sigmoid(delta + inverse_sigmoid(ref)) three times) turn into up to 0.7 px at a 1748 px crop -- above the 0.5 px bar;
  // the decoder side is always 3xTF32.  SPE_SA_X3=0 selects the plain-TF32 trunk (3x less tensor work).
  bool x3 = getenv("SPE_SA_X3") ? atoi(getenv("SPE_SA_X3")) != 0 : true;
  // SiLU / GELU inside the epilogue of the 3xTF32 GEMM kernels instead of a separate pass.  MEASURED AND LEFT OFF
  // (SPE_SA_ACT_EPI=1 enables it): 8.38 vs 7.98 ms per forward at B = 64 -- the 3xTF32 kernels spend their extra warps on
  // the operand split and keep a four-warp epilogue, which ~10 more instructions per element turn into the long pole of
  // the K <= 512 layers; the separate pass is a 3 % memory-bound tail.
  bool act_epi = getenv("SPE_SA_ACT_EPI") ? atoi(getenv("SPE_SA_ACT_EPI")) != 0 : false;
  int shapes_hw[6] = {0, 0, 0, 0, 0, 0};
  int nblocks[4] = {3, 4, 6, 3};   // per stage, read off the state_dict at load (PResNet depth 50 / 34 / 18)
  int expansion = 4;               // 4: BottleNeck, 1: BasicBlock
  GemmW c11, c12, c13;
  std::vector<SaBlock> blocks;
  GemmW eproj[3];
  GemmW a_qkv, a_out, a_ff1, a_ff2;
  float* a_addend = nullptr;
  float *an1g = nullptr, *an1b = nullptr, *an2g = nullptr, *an2b = nullptr;
  GemmW lat[2];
  SaCsp fpn[2], pan[2];
  GemmW dproj[3];
  GemmW enc_out, ebb0, ebb1, value_all, qp1;
  float *eo_g = nullptr, *eo_b = nullptr, *esc_w = nullptr, *esc_b = nullptr, *ebb2_w = nullptr, *ebb2_b = nullptr;
  float *anchors = nullptr, *qp0_w = nullptr, *qp0_b = nullptr;
  std::vector<SaDecLayer> dec;
  // workspace (fp32)
  void *IM2 = nullptr, *SA0 = nullptr, *SA1 = nullptr, *P0 = nullptr, *P1 = nullptr, *T1 = nullptr, *T2 = nullptr,
       *DS = nullptr, *AP = nullptr, *C3 = nullptr, *C4 = nullptr, *C5 = nullptr;
  void *E2 = nullptr, *QKV = nullptr, *ATT = nullptr, *X2 = nullptr, *HID = nullptr;
  void *CAT16 = nullptr, *CAT8 = nullptr, *CATP16 = nullptr, *CATP32 = nullptr, *K1 = nullptr, *K2 = nullptr,
       *K3 = nullptr, *K4 = nullptr, *I16 = nullptr, *O8 = nullptr, *O16 = nullptr, *O32 = nullptr;
  void *MEM = nullptr, *OM = nullptr, *ESC = nullptr, *EH1 = nullptr, *EH2 = nullptr, *EXY = nullptr, *TOPK = nullptr,
       *TGT = nullptr, *REFU = nullptr, *REF = nullptr, *ETL = nullptr, *VAL = nullptr;
  void *QPH = nullptr, *QPOS = nullptr, *X1 = nullptr, *DQKV = nullptr, *DATT = nullptr, *TGT2 = nullptr,
       *OFFAW = nullptr, *DHID = nullptr, *HB = nullptr, *HB2 = nullptr, *HG2 = nullptr;
  void *LOGS = nullptr, *PTS = nullptr, *SIGS = nullptr;     // [L, B, Q, *] per-layer outputs nobody asked for
  // extra outputs / inputs of the current spe_forward_sa call (null on the plain spe_forward path)
  float *aux_logits = nullptr, *aux_points = nullptr, *aux_logsig = nullptr;
  int32_t* topk_out = nullptr;
  const int32_t* topk_in = nullptr;
};

static const int kSaPlanes[4] = {64, 128, 256, 512};

// conv weight (Cout, Cin, R, R) -> [CoutPad][(r*R + s)*CinPad + c], zero rows / columns for the padding channels
static std::vector<float> sa_repack_conv_padded(const HostTensor& w, int CoutPad, int CinPad) {
  const int Cout = static_cast<int>(w.shape[0]), Cin = static_cast<int>(w.shape[1]);
  const int R = static_cast<int>(w.shape[2]), S = static_cast<int>(w.shape[3]);
  const int K = R * S * CinPad;
  std::vector<float> out(static_cast<size_t>(CoutPad) * K, 0.f);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s)
          out[static_cast<size_t>(o) * K + (r * S + s) * CinPad + c] = w.data[((static_cast<size_t>(o) * Cin + c) * R + r) * S + s];
  return out;
}

// BatchNorm2d (eval) scale / bias of `C` channels, zero for the padding channels up to Cpad
static std::string sa_load_bn_padded(spe_ctx* ctx, WeightSource& ws, const std::string& prefix, int C, int Cpad, GemmW* g) {
  std::vector<float> sc, bi;
  if (!bn_fold_host(ws, prefix, C, &sc, &bi)) return ws.missing;
  sc.resize(Cpad, 0.f);
  bi.resize(Cpad, 0.f);
  TRY_S(upload_f32(ctx, sc.data(), Cpad, &g->scale));
  TRY_S(upload_f32(ctx, bi.data(), Cpad, &g->bias));
  return "";
}

static std::string sa_load_conv_norm(spe_ctx* ctx, WeightSource& ws, const std::string& p, int Cout, int Cin, int R, GemmW* g) {
  return load_conv_bn(ctx, ws, p + ".conv", p + ".norm", Cout, Cin, R, g, 0, ctx->sa->x3);
}

// rows [N, K] (+ bias) as GEMM weights
static std::string sa_upload_linear(spe_ctx* ctx, const std::vector<float>& w, const std::vector<float>& b, int N, int K,
                                    GemmW* g, bool x3) {
  TRY_S(upload_gemm_w(ctx, w, N, K, g, x3));
  TRY_S(upload_f32(ctx, b.data(), N, &g->bias));
  return "";
}

static std::string sa_load_csp(spe_ctx* ctx, WeightSource& ws, const std::string& p, int E, int Hc, SaCsp* c) {
  TRY_S(sa_load_conv_norm(ctx, ws, p + ".conv1", Hc, 2 * E, 1, &c->c1));
  TRY_S(sa_load_conv_norm(ctx, ws, p + ".conv2", Hc, 2 * E, 1, &c->c2));
  TRY_S(sa_load_conv_norm(ctx, ws, p + ".conv3", E, Hc, 1, &c->c3));
  // RepVggBlock: conv1 (3x3) + BN and conv2 (1x1) + BN as one 3x3 convolution with bias
  const std::string r = p + ".bottlenecks.0";
  const HostTensor* w3 = ws.get(r + ".conv1.conv.weight", {Hc, Hc, 3, 3});
  const HostTensor* w1 = ws.get(r + ".conv2.conv.weight", {Hc, Hc, 1, 1});
  std::vector<float> s3, b3, s1, b1;
  if (!w3 || !w1 || !bn_fold_host(ws, r + ".conv1.norm", Hc, &s3, &b3) || !bn_fold_host(ws, r + ".conv2.norm", Hc, &s1, &b1))
    return ws.missing;
  std::vector<float> w = repack_conv(*w3);                       // [Hc][(r*3+s)*Hc + c]
  const int K = 9 * Hc;
  std::vector<float> bias(Hc);
  for (int o = 0; o < Hc; ++o) {
    float* row = w.data() + static_cast<size_t>(o) * K;
    for (int k = 0; k < K; ++k) row[k] *= s3[o];
    for (int ci = 0; ci < Hc; ++ci) row[4 * Hc + ci] += s1[o] * w1->data[static_cast<size_t>(o) * Hc + ci];   // centre tap
    bias[o] = b3[o] + b1[o];
  }
  TRY_S(upload_gemm_w(ctx, w, Hc, K, &c->rep, ctx->sa->x3));
  TRY_S(upload_f32(ctx, bias.data(), Hc, &c->rep.bias));
  return "";
}

std::string sa_load_weights(spe_ctx* ctx, WeightSource& ws) {
  SaModel& m = *ctx->sa;
  const spe_config& c = ctx->cfg;
  const int E = 256, Q = c.num_queries;
  (void)Q;
  // ---- PResNet-50-vd
  {
    const HostTensor* w1 = ws.get("backbone.conv1.conv1_1.conv.weight", {32, 3, 3, 3});
    const HostTensor* w2 = ws.get("backbone.conv1.conv1_2.conv.weight", {32, 32, 3, 3});
    const HostTensor* w3 = ws.get("backbone.conv1.conv1_3.conv.weight", {64, 32, 3, 3});
    if (!w1 || !w2 || !w3) return ws.missing;
    std::vector<float> a = repack_conv(*w1, 32);                 // [32][27 -> 32]
    a.resize(static_cast<size_t>(64) * 32, 0.f);                 // 32 zero output channels
    TRY_S(upload_gemm_w(ctx, a, 64, 32, &m.c11, m.x3));
    TRY_S(sa_load_bn_padded(ctx, ws, "backbone.conv1.conv1_1.norm", 32, 64, &m.c11));
    TRY_S(upload_gemm_w(ctx, sa_repack_conv_padded(*w2, 64, 32), 64, 9 * 32, &m.c12, m.x3));
    TRY_S(sa_load_bn_padded(ctx, ws, "backbone.conv1.conv1_2.norm", 32, 64, &m.c12));
    TRY_S(upload_gemm_w(ctx, sa_repack_conv_padded(*w3, 64, 32), 64, 9 * 32, &m.c13, m.x3));
    TRY_S(sa_load_bn_padded(ctx, ws, "backbone.conv1.conv1_3.norm", 64, 64, &m.c13));
  }
  // the recipe's depth is read off the checkpoint: BottleNeck blocks carry a branch2c (depth 50), BasicBlocks do not
  // (depth 18: 2 + 2 + 2 + 2 blocks, depth 34: 3 + 4 + 6 + 3) -- SA/nn/backbone/presnet.py:18-24, :35-123
  const bool basic = ws.t.find("backbone.res_layers.0.blocks.0.branch2c.conv.weight") == ws.t.end();
  m.expansion = basic ? 1 : 4;
  for (int si = 0; si < 4; ++si) {
    int nb = 0;
    while (ws.t.find("backbone.res_layers." + std::to_string(si) + ".blocks." + std::to_string(nb) + ".branch2a.conv.weight") != ws.t.end()) ++nb;
    if (nb < 1 || nb > 64) return "backbone.res_layers." + std::to_string(si) + ": no residual blocks found";
    m.nblocks[si] = nb;
  }
  m.blocks.clear();
  int cin = 64;
  for (int si = 0; si < 4; ++si)
    for (int bi = 0; bi < m.nblocks[si]; ++bi) {
      SaBlock bk;
      bk.cin = cin;
      bk.planes = kSaPlanes[si];
      bk.stride = (bi == 0 && si != 0) ? 2 : 1;
      bk.has_short = bi == 0;
      bk.basic = basic;
      const std::string p = "backbone.res_layers." + std::to_string(si) + ".blocks." + std::to_string(bi);
      if (basic) {
        TRY_S(sa_load_conv_norm(ctx, ws, p + ".branch2a", bk.planes, cin, 3, &bk.a));
        TRY_S(sa_load_conv_norm(ctx, ws, p + ".branch2b", bk.planes, bk.planes, 3, &bk.b));
      } else {
        TRY_S(sa_load_conv_norm(ctx, ws, p + ".branch2a", bk.planes, cin, 1, &bk.a));
        TRY_S(sa_load_conv_norm(ctx, ws, p + ".branch2b", bk.planes, bk.planes, 3, &bk.b));
        TRY_S(sa_load_conv_norm(ctx, ws, p + ".branch2c", bk.planes * 4, bk.planes, 1, &bk.c));
      }
      if (bk.has_short)
        TRY_S(sa_load_conv_norm(ctx, ws, p + (si == 0 ? ".short" : ".short.conv"), bk.planes * m.expansion, cin, 1, &bk.sc));
      cin = bk.planes * m.expansion;
      m.blocks.push_back(bk);
    }
  // ---- HybridEncoder
  const int cins[3] = {128 * m.expansion, 256 * m.expansion, 512 * m.expansion};
  for (int i = 0; i < 3; ++i) {
    const std::string p = "encoder.input_proj." + std::to_string(i);
    TRY_S(load_conv_bn(ctx, ws, p + ".0", p + ".1", E, cins[i], 1, &m.eproj[i], 0, m.x3));
  }
  {
    // 2-D sin-cos positions of the /32 level (hybrid_encoder.py:306-326): token n = y * W + x carries
    // [sin(y w), cos(y w), sin(x w), cos(x w)], w_k = 1 / 10000^(k / 64)
    const int h = m.hl[2], T = h * h, pd = E / 4;
    std::vector<float> pos(static_cast<size_t>(T) * E);
    for (int n = 0; n < T; ++n) {
      const float gi = static_cast<float>(n / h), gj = static_cast<float>(n % h);
      for (int k = 0; k < pd; ++k) {
        const float omega = 1.0f / powf(10000.0f, static_cast<float>(k) / static_cast<float>(pd));
        float* p = pos.data() + static_cast<size_t>(n) * E;
        p[k] = sinf(gi * omega); p[pd + k] = cosf(gi * omega);
        p[2 * pd + k] = sinf(gj * omega); p[3 * pd + k] = cosf(gj * omega);
      }
    }
    float* pos_dev = nullptr;
    SPE_CUDA_TRY(cudaMalloc(&pos_dev, pos.size() * sizeof(float)));
    cudaMemcpy(pos_dev, pos.data(), pos.size() * sizeof(float), cudaMemcpyHostToDevice);
    const std::string p = "encoder.encoder.0.layers.0";
    std::string s = load_mha_self(ctx, ws, p + ".self_attn", pos_dev, T, &m.a_qkv, &m.a_out, &m.a_addend, m.x3);
    cudaFree(pos_dev);
    if (!s.empty()) return s;
    TRY_S(load_linear(ctx, ws, p + ".linear1", c.dim_feedforward, E, &m.a_ff1, m.x3));
    TRY_S(load_linear(ctx, ws, p + ".linear2", E, c.dim_feedforward, &m.a_ff2, m.x3));
    TRY_S(load_vec(ctx, ws, p + ".norm1.weight", E, &m.an1g));
    TRY_S(load_vec(ctx, ws, p + ".norm1.bias", E, &m.an1b));
    TRY_S(load_vec(ctx, ws, p + ".norm2.weight", E, &m.an2g));
    TRY_S(load_vec(ctx, ws, p + ".norm2.bias", E, &m.an2b));
  }
  const int Hc = 128;   // CSPRepLayer hidden channels (expansion 0.5 of the speed configs)
  for (int i = 0; i < 2; ++i) {
    TRY_S(sa_load_conv_norm(ctx, ws, "encoder.lateral_convs." + std::to_string(i), E, E, 1, &m.lat[i]));
    TRY_S(sa_load_csp(ctx, ws, "encoder.fpn_blocks." + std::to_string(i), E, Hc, &m.fpn[i]));
    TRY_S(sa_load_csp(ctx, ws, "encoder.pan_blocks." + std::to_string(i), E, Hc, &m.pan[i]));
  }
  // ---- RTDETRTransformer
  const bool x3 = true;
  for (int i = 0; i < 3; ++i)
    TRY_S(sa_load_conv_norm(ctx, ws, "decoder.input_proj." + std::to_string(i), E, E, 1, &m.dproj[i]));
  TRY_S(load_linear(ctx, ws, "decoder.enc_output.0", E, E, &m.enc_out, x3));
  TRY_S(load_vec(ctx, ws, "decoder.enc_output.1.weight", E, &m.eo_g));
  TRY_S(load_vec(ctx, ws, "decoder.enc_output.1.bias", E, &m.eo_b));
  {
    const HostTensor* w = ws.get("decoder.enc_score_head.weight", {12, E});
    if (!w) return ws.missing;
    TRY_S(upload_f32(ctx, w->data, 12 * E, &m.esc_w));
    TRY_S(load_vec(ctx, ws, "decoder.enc_score_head.bias", 12, &m.esc_b));
    TRY_S(load_linear(ctx, ws, "decoder.enc_bbox_head.layers.0", E, E, &m.ebb0, x3));
    TRY_S(load_linear(ctx, ws, "decoder.enc_bbox_head.layers.1", E, E, &m.ebb1, x3));
    const HostTensor* w2 = ws.get("decoder.enc_bbox_head.layers.2.weight", {2, E});
    if (!w2) return ws.missing;
    TRY_S(upload_f32(ctx, w2->data, 2 * E, &m.ebb2_w));
    TRY_S(load_vec(ctx, ws, "decoder.enc_bbox_head.layers.2.bias", 2, &m.ebb2_b));
  }
  {
    // anchors: logit of the cell centres of every level (rtdetr_decoder.py:570-600); +inf outside (eps, 1 - eps)
    std::vector<float> an(static_cast<size_t>(m.Lv) * 2);
    const float eps = 1e-2f;
    for (int l = 0; l < 3; ++l) {
      const int h = m.hl[l];
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < h; ++x) {
          const float ax = (static_cast<float>(x) + 0.5f) / static_cast<float>(h);
          const float ay = (static_cast<float>(y) + 0.5f) / static_cast<float>(h);
          const bool valid = ax > eps && ax < 1.f - eps && ay > eps && ay < 1.f - eps;
          float* o = an.data() + (static_cast<size_t>(m.start[l]) + static_cast<size_t>(y) * h + x) * 2;
          o[0] = valid ? logf(ax / (1.f - ax)) : INFINITY;
          o[1] = valid ? logf(ay / (1.f - ay)) : INFINITY;
        }
    }
    TRY_S(upload_f32(ctx, an.data(), static_cast<long long>(an.size()), &m.anchors));
  }
  {
    const HostTensor* w = ws.get("decoder.query_pos_head.layers.0.weight", {2 * E, 2});
    if (!w) return ws.missing;
    TRY_S(upload_f32(ctx, w->data, 2 * E * 2, &m.qp0_w));
    TRY_S(load_vec(ctx, ws, "decoder.query_pos_head.layers.0.bias", 2 * E, &m.qp0_b));
    TRY_S(load_linear(ctx, ws, "decoder.query_pos_head.layers.1", E, 2 * E, &m.qp1, x3));
  }
  const int LD = c.dec_layers;
  m.dec.assign(LD, SaDecLayer{});
  std::vector<float> vw(static_cast<size_t>(LD) * E * E), vb(static_cast<size_t>(LD) * E);
  for (int i = 0; i < LD; ++i) {
    SaDecLayer& L = m.dec[i];
    const std::string p = "decoder.decoder.layers." + std::to_string(i);
    const HostTensor* w = ws.get(p + ".self_attn.in_proj_weight", {3 * E, E});
    const HostTensor* b = ws.get(p + ".self_attn.in_proj_bias", {3 * E});
    if (!w || !b) return ws.missing;
    TRY_S(sa_upload_linear(ctx, std::vector<float>(w->data, w->data + 2 * E * E), std::vector<float>(b->data, b->data + 2 * E),
                           2 * E, E, &L.sa_qk, x3));
    TRY_S(sa_upload_linear(ctx, std::vector<float>(w->data + 2 * E * E, w->data + 3 * E * E),
                           std::vector<float>(b->data + 2 * E, b->data + 3 * E), E, E, &L.sa_v, x3));
    TRY_S(load_linear(ctx, ws, p + ".self_attn.out_proj", E, E, &L.sa_out, x3));
    const int NO = 8 * 3 * 4 * 2, NA = 8 * 3 * 4, NP = 320;
    const HostTensor* wo = ws.get(p + ".cross_attn.sampling_offsets.weight", {NO, E});
    const HostTensor* bo = ws.get(p + ".cross_attn.sampling_offsets.bias", {NO});
    const HostTensor* wa = ws.get(p + ".cross_attn.attention_weights.weight", {NA, E});
    const HostTensor* ba = ws.get(p + ".cross_attn.attention_weights.bias", {NA});
    const HostTensor* wv = ws.get(p + ".cross_attn.value_proj.weight", {E, E});
    const HostTensor* bv = ws.get(p + ".cross_attn.value_proj.bias", {E});
    if (!wo || !bo || !wa || !ba || !wv || !bv) return ws.missing;
    std::vector<float> ow(static_cast<size_t>(NP) * E, 0.f), ob(NP, 0.f);
    memcpy(ow.data(), wo->data, sizeof(float) * NO * E);
    memcpy(ow.data() + static_cast<size_t>(NO) * E, wa->data, sizeof(float) * NA * E);
    memcpy(ob.data(), bo->data, sizeof(float) * NO);
    memcpy(ob.data() + NO, ba->data, sizeof(float) * NA);
    TRY_S(sa_upload_linear(ctx, ow, ob, NP, E, &L.offaw, x3));
    memcpy(vw.data() + static_cast<size_t>(i) * E * E, wv->data, sizeof(float) * E * E);
    memcpy(vb.data() + static_cast<size_t>(i) * E, bv->data, sizeof(float) * E);
    TRY_S(load_linear(ctx, ws, p + ".cross_attn.output_proj", E, E, &L.ca_out, x3));
    TRY_S(load_linear(ctx, ws, p + ".linear1", c.dim_feedforward, E, &L.ff1, x3));
    TRY_S(load_linear(ctx, ws, p + ".linear2", E, c.dim_feedforward, &L.ff2, x3));
    TRY_S(load_vec(ctx, ws, p + ".norm1.weight", E, &L.n1g));
    TRY_S(load_vec(ctx, ws, p + ".norm1.bias", E, &L.n1b));
    TRY_S(load_vec(ctx, ws, p + ".norm2.weight", E, &L.n2g));
    TRY_S(load_vec(ctx, ws, p + ".norm2.bias", E, &L.n2b));
    TRY_S(load_vec(ctx, ws, p + ".norm3.weight", E, &L.n3g));
    TRY_S(load_vec(ctx, ws, p + ".norm3.bias", E, &L.n3b));
    // heads of this layer
    const std::string is = std::to_string(i);
    const HostTensor* b0w = ws.get("decoder.dec_bbox_head." + is + ".layers.0.weight", {E, E});
    const HostTensor* b0b = ws.get("decoder.dec_bbox_head." + is + ".layers.0.bias", {E});
    const HostTensor* s0w = ws.get("decoder.decoder.sigma_embed." + is + ".layers.0.weight", {E, E});
    const HostTensor* s0b = ws.get("decoder.decoder.sigma_embed." + is + ".layers.0.bias", {E});
    if (!b0w || !b0b || !s0w || !s0b) return ws.missing;
    std::vector<float> hw(static_cast<size_t>(2) * E * E), hb(2 * E);
    memcpy(hw.data(), b0w->data, sizeof(float) * E * E);
    memcpy(hw.data() + static_cast<size_t>(E) * E, s0w->data, sizeof(float) * E * E);
    memcpy(hb.data(), b0b->data, sizeof(float) * E);
    memcpy(hb.data() + E, s0b->data, sizeof(float) * E);
    TRY_S(sa_upload_linear(ctx, hw, hb, 2 * E, E, &L.hb0, x3));
    TRY_S(load_linear(ctx, ws, "decoder.dec_bbox_head." + is + ".layers.1", E, E, &L.bb1, x3));
    TRY_S(load_linear(ctx, ws, "decoder.decoder.sigma_embed." + is + ".layers.1", E, E, &L.sg1, x3));
    const HostTensor* cw = ws.get("decoder.dec_score_head." + is + ".weight", {12, E});
    const HostTensor* b2 = ws.get("decoder.dec_bbox_head." + is + ".layers.2.weight", {2, E});
    const HostTensor* s2 = ws.get("decoder.decoder.sigma_embed." + is + ".layers.2.weight", {1, E});
    if (!cw || !b2 || !s2) return ws.missing;
    TRY_S(upload_f32(ctx, cw->data, 12 * E, &L.cls_w));
    TRY_S(load_vec(ctx, ws, "decoder.dec_score_head." + is + ".bias", 12, &L.cls_b));
    TRY_S(upload_f32(ctx, b2->data, 2 * E, &L.bb2_w));
    TRY_S(load_vec(ctx, ws, "decoder.dec_bbox_head." + is + ".layers.2.bias", 2, &L.bb2_b));
    TRY_S(upload_f32(ctx, s2->data, E, &L.sg2_w));
    TRY_S(load_vec(ctx, ws, "decoder.decoder.sigma_embed." + is + ".layers.2.bias", 1, &L.sg2_b));
  }
  TRY_S(sa_upload_linear(ctx, vw, vb, LD * E, E, &m.value_all, x3));
  return "";
}

std::string sa_alloc_workspace(spe_ctx* ctx) {
  const spe_config& c = ctx->cfg;
  ctx->sa = new SaModel();
  SaModel& m = *ctx->sa;
  m.R = c.input_size;
  int start = 0;
  for (int l = 0; l < 3; ++l) {
    m.hl[l] = c.input_size / (8 << l);
    m.start[l] = start;
    m.shapes_hw[2 * l] = m.shapes_hw[2 * l + 1] = m.hl[l];
    start += m.hl[l] * m.hl[l];
  }
  m.Lv = start;
  const long long B = c.max_batch, R = c.input_size;
  const long long h2 = R / 2, h4 = R / 4, h8 = R / 8, h16 = R / 16, h32 = R / 32;
  const long long Q = c.num_queries, LD = c.dec_layers, FF = c.dim_feedforward, Lv = m.Lv;
  auto A = [&](void** p, long long elems) {
    std::string e = dmalloc_bytes(ctx, p, elems * 4);
    if (e.empty()) ctx->ws_fields.push_back({p, elems * 4});
    return e;
  };
  TRY_S(A(&m.IM2, B * h2 * h2 * 32));
  TRY_S(A(&m.SA0, B * h2 * h2 * 64));
  TRY_S(A(&m.SA1, B * h2 * h2 * 64));
  TRY_S(A(&m.P0, B * h4 * h4 * 256));
  TRY_S(A(&m.P1, B * h4 * h4 * 256));
  TRY_S(A(&m.T1, B * h4 * h4 * 128));
  TRY_S(A(&m.T2, B * h4 * h4 * 128));
  TRY_S(A(&m.DS, B * h4 * h4 * 256));
  TRY_S(A(&m.AP, B * h8 * h8 * 256));
  TRY_S(A(&m.C3, B * h8 * h8 * 512));
  TRY_S(A(&m.C4, B * h16 * h16 * 1024));
  TRY_S(A(&m.C5, B * h32 * h32 * 2048));
  const long long T5 = h32 * h32;
  TRY_S(A(&m.E2, B * T5 * 256));
  TRY_S(A(&m.QKV, B * T5 * 768));
  TRY_S(A(&m.ATT, B * T5 * 256));
  TRY_S(A(&m.X2, B * T5 * 256));
  TRY_S(A(&m.HID, B * T5 * FF));
  TRY_S(A(&m.CAT16, B * h16 * h16 * 512));
  TRY_S(A(&m.CAT8, B * h8 * h8 * 512));
  TRY_S(A(&m.CATP16, B * h16 * h16 * 512));
  TRY_S(A(&m.CATP32, B * h32 * h32 * 512));
  TRY_S(A(&m.K1, B * h8 * h8 * 128));
  TRY_S(A(&m.K2, B * h8 * h8 * 128));
  TRY_S(A(&m.K3, B * h8 * h8 * 128));
  TRY_S(A(&m.K4, B * h8 * h8 * 256));
  TRY_S(A(&m.I16, B * h16 * h16 * 256));
  TRY_S(A(&m.O8, B * h8 * h8 * 256));
  TRY_S(A(&m.O16, B * h16 * h16 * 256));
  TRY_S(A(&m.O32, B * h32 * h32 * 256));
  TRY_S(A(&m.MEM, B * Lv * 256));
  TRY_S(A(&m.OM, B * Lv * 256));
  TRY_S(A(&m.ESC, B * Lv * 12));
  TRY_S(A(&m.EH1, B * Lv * 256));
  TRY_S(A(&m.EH2, B * Lv * 256));
  TRY_S(A(&m.EXY, B * Lv * 2));
  TRY_S(A(&m.TOPK, B * Q));
  TRY_S(A(&m.TGT, B * Q * 256));
  TRY_S(A(&m.REFU, B * Q * 2));
  TRY_S(A(&m.REF, B * Q * 2));
  TRY_S(A(&m.ETL, B * Q * 12));
  TRY_S(A(&m.VAL, B * Lv * LD * 256));
  TRY_S(A(&m.QPH, B * Q * 512));
  TRY_S(A(&m.QPOS, B * Q * 256));
  TRY_S(A(&m.X1, B * Q * 256));
  TRY_S(A(&m.DQKV, B * Q * 768));
  TRY_S(A(&m.DATT, B * Q * 256));
  TRY_S(A(&m.TGT2, B * Q * 256));
  TRY_S(A(&m.OFFAW, B * Q * 320));
  TRY_S(A(&m.DHID, B * Q * FF));
  TRY_S(A(&m.HB, B * Q * 512));
  TRY_S(A(&m.HB2, B * Q * 256));
  TRY_S(A(&m.HG2, B * Q * 256));
  TRY_S(A(&m.LOGS, LD * B * Q * 12));
  TRY_S(A(&m.PTS, LD * B * Q * 2));
  TRY_S(A(&m.SIGS, LD * B * Q * 2));
  return "";
}

namespace {

struct SaFwd {
  spe_ctx* ctx;
  SaModel& m;
  Fwd f;
  cudaStream_t st;
  long long B;

  static float* F(void* p) { return static_cast<float*>(p); }
  static float* col(void* base, long long off) { return static_cast<float*>(base) + off; }

  // out = A W^T (+ bias, BN scale) (+ residual) with an explicit A row stride; exact: leave the result unrounded
  std::string gemm(const void* A, int lda, long long M, const GemmW& w, void* out, int out_ld, bool relu, bool exact,
                   const void* residual = nullptr, int res_ld = 0, int res_mod = 0, int res_f32 = 0, int act = 0) {
    if (lda == w.K) TRY_S(f.calibrate_layer(A, M, w.K, w.K, w));
    GemmDesc d;
    d.mode = 0;
    d.A = A; d.M = M; d.K = w.K; d.lda = lda;
    d.Wt = w.w; d.N = w.N;
    d.scale = w.scale; d.bias = w.bias;
    d.residual = residual; d.res_ld = res_ld; d.res_mod = res_mod; d.res_f32 = res_f32;
    d.relu = act > 1 ? act : (relu ? 1 : 0);     // act 2 / 3: SiLU / GELU in the epilogue (3xTF32 kernels only)
    d.out = out; d.out_ld = out_ld;
    d.x3 = w.x3;
    d.round_out = (w.x3 || exact) ? 0 : 1;
    return launch_gemm(kTF32, d, ctx->num_sms, st);
  }
  // 3x3 / stride 1 convolution + BN + residual + ReLU (second convolution of a BasicBlock), H = map extent
  std::string conv_res(const void* x, int H, int C, const GemmW& w, void* out, int out_ld, const void* residual) {
    TRY_S(f.calibrate_layer(x, B * H * H, C, C, w));
    GemmDesc d;
    d.mode = 1;
    d.A = x; d.NB = static_cast<int>(B); d.H = H; d.W = H; d.C = C; d.R = 3; d.S = 3; d.pad = 1; d.conv_stride = 1;
    d.Wt = w.w; d.N = w.N;
    d.scale = w.scale; d.bias = w.bias;
    d.residual = residual; d.res_ld = out_ld;
    d.relu = 1;
    d.out = out; d.out_ld = out_ld;
    d.x3 = w.x3;
    d.round_out = w.x3 ? 0 : 1;
    return launch_gemm(kTF32, d, ctx->num_sms, st);
  }
  std::string act(const void* in, int in_ld, const void* add, int add_ld, void* out, int out_ld, long long rows, int C,
                  int kind, int round = -1) {
    if (round < 0) round = m.x3 ? 0 : 1;   // a 3xTF32 consumer splits its operand itself
    return launch_act_rows(F(const_cast<void*>(in)), in_ld, add ? F(const_cast<void*>(add)) : nullptr, add_ld, F(out),
                           out_ld, rows, C, kind, round, st);
  }
  // ConvNormLayer 1x1 + SiLU: x [rows, w.K] -> out[:, 0:w.N] (row stride out_ld)
  std::string conv1_silu(const void* x, long long rows, const GemmW& w, void* tmp, void* out, int out_ld) {
    if (w.x3 && m.act_epi) return gemm(x, w.K, rows, w, out, out_ld, false, true, nullptr, 0, 0, 0, 2);   // SiLU in the epilogue
    TRY_S(gemm(x, w.K, rows, w, tmp, w.N, false, true));
    return act(tmp, w.N, nullptr, 0, out, out_ld, rows, w.N, 1);
  }
  // CSPRepLayer (hybrid_encoder.py:96-123): x [rows, 512] at H x H -> out [rows, 256]
  std::string csp(const void* x, int H, const SaCsp& c, void* out) {
    const long long rows = B * H * H;
    TRY_S(conv1_silu(x, rows, c.c1, m.K1, m.K1, 128));
    TRY_S(conv1_silu(x, rows, c.c2, m.K2, m.K2, 128));
    TRY_S(f.conv(m.K1, H, 128, 3, 1, c.rep, m.K3, 128, false, true));
    TRY_S(act(m.K3, 128, m.K2, 128, m.K3, 128, rows, 128, 1));          // silu(rep(x1)) + x2
    return conv1_silu(m.K3, rows, c.c3, m.K4, out, 256);
  }
};

}  // namespace

std::string sa_forward(spe_ctx* ctx, const float* images, int Bi, float* logits, float* points, float* logsig,
                       cudaStream_t st) {
  if (ctx->dt != kTF32) return "the SA predictor is built for fp32 storage (TF32 / 3xTF32 tensor cores) only";
  SaModel& m = *ctx->sa;
  const spe_config& c = ctx->cfg;
  SaFwd s{ctx, m, Fwd{ctx, st, Bi, kTF32, 4}, st, Bi};
  Fwd& f = s.f;
  const long long B = Bi;
  const int R = c.input_size, h2 = R / 2, h4 = R / 4;
  const int Q = c.num_queries, LD = c.dec_layers, FF = c.dim_feedforward;
  const long long MQ = B * Q;

  // ---- PResNet stem: conv1_1 (3x3 / s2) as im2col + GEMM, conv1_2, conv1_3 (3x3), max-pool
  const int rnd = m.x3 ? 0 : 1;
  TRY_S(launch_sa_stem_im2col(images, Bi, R, R, SaFwd::F(m.IM2), rnd, st));
  TRY_S(s.gemm(m.IM2, 32, B * h2 * h2, m.c11, m.SA0, 64, true, false));
  // conv1_1 / conv1_2 write 64-channel rows whose upper half is zero; their consumers read channels [0, 32) only
  TRY_S(f.conv(m.SA0, h2, 32, 3, 1, m.c12, m.SA1, 64, true, false, 64));
  TRY_S(f.conv(m.SA1, h2, 32, 3, 1, m.c13, m.SA0, 64, true, false, 64));
  TRY_S(f.tap("sa_stem", m.SA0, B * h2 * h2 * 64));
  TRY_S(launch_maxpool3x3s2(kTF32, m.SA0, Bi, h2, h2, 64, m.P0, st));

  // ---- residual stages
  const void* cur = m.P0;
  int H = h4, bidx = 0;
  void* stage_out[4] = {nullptr, m.C3, m.C4, m.C5};
  for (int si = 0; si < 4; ++si) {
    for (int bi = 0; bi < m.nblocks[si]; ++bi, ++bidx) {
      const SaBlock& bk = m.blocks[bidx];
      const int Ho = H / bk.stride;
      const long long Min = B * H * H, Mout = B * Ho * Ho;
      const int cout = bk.planes * m.expansion;
      void* nxt = (bi == m.nblocks[si] - 1 && stage_out[si] != nullptr) ? stage_out[si] : ((cur == m.P0) ? m.P1 : m.P0);
      if (bk.basic) TRY_S(f.conv(cur, H, bk.cin, 3, bk.stride, bk.a, m.T1, bk.planes, true));
      else {
        TRY_S(s.gemm(cur, bk.cin, Min, bk.a, m.T1, bk.planes, true, false));
        TRY_S(f.conv(m.T1, H, bk.planes, 3, bk.stride, bk.b, m.T2, bk.planes, true));
      }
      const void* identity = cur;
      if (bk.has_short) {
        const void* src = cur;
        if (bk.stride == 2) {                                             // variant d: AvgPool2d(2, 2) then 1x1
          TRY_S(launch_avgpool2x2(SaFwd::F(const_cast<void*>(cur)), Bi, H, H, bk.cin, SaFwd::F(m.AP), rnd, st));
          src = m.AP;
        }
        TRY_S(s.gemm(src, bk.cin, Mout, bk.sc, m.DS, cout, false, true));
        identity = m.DS;
      }
      if (bk.basic) TRY_S(s.conv_res(m.T1, Ho, bk.planes, bk.b, nxt, cout, identity));   // 3x3 + BN, + shortcut, ReLU
      else TRY_S(s.gemm(m.T2, bk.planes, Mout, bk.c, nxt, cout, true, false, identity, cout));
      cur = nxt;
      H = Ho;
    }
    const std::string nm = "sa_stage" + std::to_string(si);
    TRY_S(f.tap(nm.c_str(), cur, B * H * H * kSaPlanes[si] * m.expansion));
  }

  // ---- HybridEncoder: projections (levels 0 / 1 straight into their concatenation slots)
  const int h8 = m.hl[0], h16 = m.hl[1], h32 = m.hl[2];
  const long long M8 = B * h8 * h8, M16 = B * h16 * h16, M32 = B * h32 * h32;
  TRY_S(s.gemm(m.C3, m.eproj[0].K, M8, m.eproj[0], SaFwd::col(m.CAT8, 256), 512, false, false));
  TRY_S(s.gemm(m.C4, m.eproj[1].K, M16, m.eproj[1], SaFwd::col(m.CAT16, 256), 512, false, false));
  TRY_S(s.gemm(m.C5, m.eproj[2].K, M32, m.eproj[2], m.E2, 256, false, false));
  // AIFI layer on the /32 level (post-norm, GELU)
  {
    const int T = h32 * h32;
    TRY_S(s.gemm(m.E2, 256, M32, m.a_qkv, m.QKV, 768, false, false, m.a_addend, 768, T, 1));
    TRY_S(f.attn(m.QKV, 768, SaFwd::col(m.QKV, 256), 768, SaFwd::col(m.QKV, 512), 768, m.ATT, T, T, m.x3 ? 1 : 0, 0, m.x3 ? 1 : 0));
    TRY_S(s.gemm(m.ATT, 256, M32, m.a_out, m.X2, 256, false, true, m.E2, 256));
    TRY_S(f.ln(m.X2, m.an1g, m.an1b, M32, m.E2, m.x3 ? 1 : 0));
    if (m.a_ff1.x3 && m.act_epi) {
      TRY_S(s.gemm(m.E2, 256, M32, m.a_ff1, m.HID, FF, false, true, nullptr, 0, 0, 0, 3));   // GELU in the epilogue
    } else {
      TRY_S(s.gemm(m.E2, 256, M32, m.a_ff1, m.HID, FF, false, true));
      TRY_S(s.act(m.HID, FF, nullptr, 0, m.HID, FF, M32, FF, 2));
    }
    TRY_S(s.gemm(m.HID, FF, M32, m.a_ff2, m.X2, 256, false, true, m.E2, 256));
    TRY_S(f.ln(m.X2, m.an2g, m.an2b, M32, m.E2, m.x3 ? 1 : 0));
    TRY_S(f.tap("sa_aifi", m.E2, M32 * 256));
  }
  // top-down: lateral conv -> nearest x2 -> concat with the finer level -> CSPRep
  TRY_S(s.conv1_silu(m.E2, M32, m.lat[0], m.K4, SaFwd::col(m.CATP32, 256), 512));           // inner_outs[2]
  TRY_S(launch_upsample_nearest2x(SaFwd::col(m.CATP32, 256), 512, Bi, h32, h32, 256, SaFwd::F(m.CAT16), 512, st));
  TRY_S(s.csp(m.CAT16, h16, m.fpn[0], m.I16));
  TRY_S(s.conv1_silu(m.I16, M16, m.lat[1], m.K4, SaFwd::col(m.CATP16, 256), 512));          // inner_outs[1]
  TRY_S(launch_upsample_nearest2x(SaFwd::col(m.CATP16, 256), 512, Bi, h16, h16, 256, SaFwd::F(m.CAT8), 512, st));
  TRY_S(s.csp(m.CAT8, h8, m.fpn[1], m.O8));                                                  // outs[0]
  // bottom-up: bicubic x0.5 -> concat with the lateral output -> CSPRep
  TRY_S(launch_bicubic_half(SaFwd::F(m.O8), Bi, h8, h8, 256, SaFwd::F(m.CATP16), 512, rnd, st));
  TRY_S(s.csp(m.CATP16, h16, m.pan[0], m.O16));                                              // outs[1]
  TRY_S(launch_bicubic_half(SaFwd::F(m.O16), Bi, h16, h16, 256, SaFwd::F(m.CATP32), 512, rnd, st));
  TRY_S(s.csp(m.CATP32, h32, m.pan[1], m.O32));                                              // outs[2]
  TRY_S(f.tap("sa_enc0", m.O8, M8 * 256));
  TRY_S(f.tap("sa_enc1", m.O16, M16 * 256));
  TRY_S(f.tap("sa_enc2", m.O32, M32 * 256));

  // ---- decoder input: memory [B, Lv, 256] = concat over levels of BN(conv1x1(out_l))
  const long long Lv = m.Lv;
  {
    const void* outs[3] = {m.O8, m.O16, m.O32};
    const long long Ml[3] = {M8, M16, M32};
    for (int l = 0; l < 3; ++l) {
      TRY_S(s.gemm(outs[l], 256, Ml[l], m.dproj[l], m.K4, 256, false, true));
      const size_t w = static_cast<size_t>(m.hl[l]) * m.hl[l] * 256 * 4;
      SPE_CUDA_TRY(cudaMemcpy2DAsync(SaFwd::col(m.MEM, static_cast<long long>(m.start[l]) * 256), static_cast<size_t>(Lv) * 256 * 4,
                                     m.K4, w, w, static_cast<size_t>(B), cudaMemcpyDeviceToDevice, st));
    }
  }
  TRY_S(f.tap("sa_memory", m.MEM, B * Lv * 256));
  // enc_output (Linear + LayerNorm), class scores and keypoint logits of every anchor
  TRY_S(s.gemm(m.MEM, 256, B * Lv, m.enc_out, m.EH1, 256, false, true));
  TRY_S(f.ln(m.EH1, m.eo_g, m.eo_b, B * Lv, m.OM, 1));
  TRY_S(launch_small_linear(SaFwd::F(m.OM), 256, B * Lv, 256, m.esc_w, m.esc_b, 12, SaFwd::F(m.ESC), 12, nullptr, 0, st));
  TRY_S(s.gemm(m.OM, 256, B * Lv, m.ebb0, m.EH1, 256, true, true));
  TRY_S(s.gemm(m.EH1, 256, B * Lv, m.ebb1, m.EH2, 256, true, true));
  TRY_S(launch_small_linear(SaFwd::F(m.EH2), 256, B * Lv, 256, m.ebb2_w, m.ebb2_b, 2, SaFwd::F(m.EXY), 2, m.anchors,
                            static_cast<int>(Lv), st));
  TRY_S(f.tap("sa_enc_scores", m.ESC, B * Lv * 12));
  // top-k query selection
  int32_t* topk = static_cast<int32_t*>(m.TOPK);
  if (m.topk_in != nullptr) {
    SPE_CUDA_TRY(cudaMemcpyAsync(topk, m.topk_in, static_cast<size_t>(MQ) * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    TRY_S(launch_topk_queries(SaFwd::F(m.ESC), Bi, static_cast<int>(Lv), 12, Q, topk, nullptr, st));
  }
  if (m.topk_out != nullptr)
    SPE_CUDA_TRY(cudaMemcpyAsync(m.topk_out, topk, static_cast<size_t>(MQ) * 4, cudaMemcpyDeviceToDevice, st));
  TRY_S(launch_gather_rows(SaFwd::F(m.OM), topk, Bi, static_cast<int>(Lv), Q, 256, SaFwd::F(m.TGT), st));
  TRY_S(launch_gather_rows(SaFwd::F(m.EXY), topk, Bi, static_cast<int>(Lv), Q, 2, SaFwd::F(m.REFU), st));
  float* enc_logits = m.aux_logits ? m.aux_logits + static_cast<long long>(LD - 1) * MQ * 12 : SaFwd::F(m.ETL);
  TRY_S(launch_gather_rows(SaFwd::F(m.ESC), topk, Bi, static_cast<int>(Lv), Q, 12, enc_logits, st));
  TRY_S(s.act(m.REFU, 4, nullptr, 0, m.REF, 4, MQ * 2 / 4, 4, 3, 0));                       // sigmoid
  if (m.aux_points)
    SPE_CUDA_TRY(cudaMemcpyAsync(m.aux_points + static_cast<long long>(LD - 1) * MQ * 2, m.REF, static_cast<size_t>(MQ) * 2 * 4,
                                 cudaMemcpyDeviceToDevice, st));
  // value projections of all decoder layers
  TRY_S(s.gemm(m.MEM, 256, B * Lv, m.value_all, m.VAL, LD * 256, false, true));

  // ---- decoder layers
  for (int i = 0; i < LD; ++i) {
    const SaDecLayer& L = m.dec[i];
    TRY_S(launch_query_pos_hidden(SaFwd::F(m.REF), m.qp0_w, m.qp0_b, 512, MQ, SaFwd::F(m.QPH), st));
    TRY_S(s.gemm(m.QPH, 512, MQ, m.qp1, m.QPOS, 256, false, true));
    // self-attention: q = k = tgt + query_pos, v = tgt
    TRY_S(launch_add(SaFwd::F(m.TGT), SaFwd::F(m.QPOS), SaFwd::F(m.X1), MQ * 256, st));
    TRY_S(s.gemm(m.X1, 256, MQ, L.sa_qk, m.DQKV, 768, false, true));
    TRY_S(s.gemm(m.TGT, 256, MQ, L.sa_v, SaFwd::col(m.DQKV, 512), 768, false, true));
    TRY_S(f.attn(m.DQKV, 768, SaFwd::col(m.DQKV, 256), 768, SaFwd::col(m.DQKV, 512), 768, m.DATT, Q, Q, 1, 0, 1));
    TRY_S(s.gemm(m.DATT, 256, MQ, L.sa_out, m.TGT2, 256, false, true, m.TGT, 256));
    TRY_S(f.ln(m.TGT2, L.n1g, L.n1b, MQ, m.TGT, 1));
    // multi-scale deformable cross-attention around the reference points
    TRY_S(launch_add(SaFwd::F(m.TGT), SaFwd::F(m.QPOS), SaFwd::F(m.X1), MQ * 256, st));
    TRY_S(s.gemm(m.X1, 256, MQ, L.offaw, m.OFFAW, 320, false, true));
    TRY_S(launch_ms_deform_attn(SaFwd::col(m.VAL, static_cast<long long>(i) * 256), m.shapes_hw, 3, SaFwd::F(m.OFFAW),
                                SaFwd::col(m.OFFAW, 192), SaFwd::F(m.REF), 1, Bi, Q, 8, 4, 1, SaFwd::F(m.DATT), st,
                                static_cast<long long>(LD) * 256, 320, 320));
    TRY_S(s.gemm(m.DATT, 256, MQ, L.ca_out, m.TGT2, 256, false, true, m.TGT, 256));
    TRY_S(f.ln(m.TGT2, L.n2g, L.n2b, MQ, m.TGT, 1));
    // feed-forward
    TRY_S(s.gemm(m.TGT, 256, MQ, L.ff1, m.DHID, FF, true, true));
    TRY_S(s.gemm(m.DHID, FF, MQ, L.ff2, m.TGT2, 256, false, true, m.TGT, 256));
    TRY_S(f.ln(m.TGT2, L.n3g, L.n3b, MQ, m.TGT, 1));
    const std::string nm = "sa_dec" + std::to_string(i);
    TRY_S(f.tap(nm.c_str(), m.TGT, MQ * 256));
    // heads: class logits, refined keypoints, log-sigma
    TRY_S(s.gemm(m.TGT, 256, MQ, L.hb0, m.HB, 512, true, true));
    TRY_S(s.gemm(m.HB, 512, MQ, L.bb1, m.HB2, 256, true, true));
    TRY_S(s.gemm(SaFwd::col(m.HB, 256), 512, MQ, L.sg1, m.HG2, 256, true, true));
    const bool last = i == LD - 1;
    float* lo = last ? logits : (m.aux_logits ? m.aux_logits + i * MQ * 12 : SaFwd::F(m.LOGS) + i * MQ * 12);
    float* po = last ? points : (m.aux_points ? m.aux_points + i * MQ * 2 : SaFwd::F(m.PTS) + i * MQ * 2);
    float* so = last ? (logsig ? logsig : SaFwd::F(m.SIGS) + i * MQ * 2)
                     : (m.aux_logsig ? m.aux_logsig + i * MQ * 2 : SaFwd::F(m.SIGS) + i * MQ * 2);
    TRY_S(launch_sa_head(SaFwd::F(m.TGT), SaFwd::F(m.HB2), SaFwd::F(m.HG2), SaFwd::F(m.REF), MQ, L.cls_w, L.cls_b,
                         L.bb2_w, L.bb2_b, L.sg2_w, L.sg2_b, lo, po, so, SaFwd::F(m.REF), st));
  }
  return "";
}
