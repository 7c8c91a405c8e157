// ROS-free harness: drives the C++ Cerebro shim (cerebro_shim.hpp) over libcerebro_b200.so exactly the way
// cerebro_node.cpp:487-509 drives the three threads, on keyframes read from a raw file.
//
//   harness <weights.cbw> <images.raw> <n> <rows> <cols> <chnls> [<pnp.raw> <npts>]
//   harness --consistency <trials.raw> <n>     (no device call: ProcessedLoopCandidate::makeLoopEdgeMsgWithConsistencyCheck on n
//                                               trials of 54 doubles: t_1, t_2 [s], pf_matches, 3 goodness values, op1, op2, icp
//                                               row-major 4x4; prints one JSON object per trial)
//
// images.raw : n * rows*cols*chnls bytes; keyframes arrive 3 at a time (the search thread acts on >= 3 new
// descriptors, src/Cerebro.cpp:962).  pnp.raw : npts*3 doubles (w_X) followed by npts*2 doubles (uv).
// Prints one JSON object: {"descriptor_size":..,"found":[[curr,prev,score],..],"pnp":{"confidence":..,"T":[16]}}
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "cerebro_shim.hpp"

namespace {

struct Cbw {
  std::vector<std::string> names;
  std::vector<std::vector<int>> shapes;
  std::vector<std::vector<float>> data;
  std::vector<int> strides;
  std::vector<int> residual;  // MobileNetV2 containers: Add flag per inverted-residual block
  bool v2 = false;            // header "arch": "mobilenetv2"
  const std::vector<float>* get(const std::string& n, std::vector<int>* shape = nullptr) const {
    for (size_t i = 0; i < names.size(); ++i)
      if (names[i] == n) {
        if (shape) *shape = shapes[i];
        return &data[i];
      }
    return nullptr;
  }
};

std::vector<int> parse_int_list(const std::string& s, size_t& p) {  // p at '['
  std::vector<int> out;
  ++p;
  while (p < s.size() && s[p] != ']') {
    while (p < s.size() && (s[p] == ' ' || s[p] == ',')) ++p;
    if (s[p] == ']') break;
    out.push_back(std::atoi(s.c_str() + p));
    while (p < s.size() && s[p] != ',' && s[p] != ']') ++p;
  }
  ++p;
  return out;
}

bool load_cbw(const char* path, Cbw& w) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  char magic[4];
  uint32_t hl = 0;
  f.read(magic, 4);
  f.read(reinterpret_cast<char*>(&hl), 4);
  if (std::memcmp(magic, "CBW1", 4) != 0) return false;
  std::string h(hl, '\0');
  f.read(&h[0], hl);
  w.v2 = h.find("\"arch\": \"mobilenetv2\"") != std::string::npos;
  size_t p = h.find("\"strides\"");
  p = h.find('[', p);
  w.strides = parse_int_list(h, p);
  if (w.v2) {
    p = h.find("\"residual\"");
    p = h.find('[', p);
    w.residual = parse_int_list(h, p);
  }
  p = h.find("\"arrays\"");
  p = h.find('[', p) + 1;  // inside the outer list
  while (true) {
    size_t q = h.find("[\"", p);
    if (q == std::string::npos) break;
    size_t e = h.find('"', q + 2);
    std::string name = h.substr(q + 2, e - q - 2);
    if (name == "meta") break;
    size_t b = h.find('[', e);
    std::vector<int> shape = parse_int_list(h, b);
    w.names.push_back(name);
    w.shapes.push_back(shape);
    p = b;
    if (h.compare(p, 3, "]],") != 0 && h.find("[\"", p) > h.find("\"meta\"", p)) break;
  }
  for (size_t i = 0; i < w.names.size(); ++i) {
    size_t cnt = 1;
    for (int d : w.shapes[i]) cnt *= (size_t)d;
    std::vector<float> a(cnt);
    f.read(reinterpret_cast<char*>(a.data()), (std::streamsize)(cnt * sizeof(float)));
    w.data.push_back(std::move(a));
  }
  return (bool)f;
}

}  // namespace

int run_consistency(const char* path, int n) {
  using namespace cerebro_b200;
  std::ifstream f(path, std::ios::binary);
  if (!f) {
    std::fprintf(stderr, "cannot read %s\n", path);
    return 2;
  }
  std::vector<double> v(54);
  for (int i = 0; i < n; ++i) {
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(54 * sizeof(double)));
    if (!f) return 2;
    Time t1, t2;
    t1.nsec = (int64_t)std::llround(v[0] * 1e9);
    t2.nsec = (int64_t)std::llround(v[1] * 1e9);
    ProcessedLoopCandidate c(i, t1, t2, 2 * i, 2 * i + 1);
    c.pf_matches = (int)v[2];
    c.opX_goodness = {(float)v[3], (float)v[4], (float)v[5]};
    c.opX_b_T_a = {Matrix4d::from(&v[6]), Matrix4d::from(&v[22]), Matrix4d::from(&v[38])};
    LoopEdge e;
    const bool ok = c.makeLoopEdgeMsgWithConsistencyCheck(e);
    std::printf("{\"ok\":%d,\"weight\":%.9g,\"position\":[%.17g,%.17g,%.17g],\"orientation\":[%.17g,%.17g,%.17g,%.17g],\"description\":\"%s\"}\n",
                ok ? 1 : 0, ok ? (double)e.weight : 0.0, e.position[0], e.position[1], e.position[2], e.orientation_xyzw[0],
                e.orientation_xyzw[1], e.orientation_xyzw[2], e.orientation_xyzw[3], ok ? e.description.c_str() : "");
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && std::strcmp(argv[1], "--consistency") == 0) return run_consistency(argv[2], std::atoi(argv[3]));
  if (argc < 7) {
    std::fprintf(stderr, "usage: harness weights.cbw images.raw n rows cols chnls [pnp.raw npts]\n");
    return 2;
  }
  Cbw w;
  if (!load_cbw(argv[1], w)) {
    std::fprintf(stderr, "cannot read %s\n", argv[1]);
    return 2;
  }
  const int n = std::atoi(argv[3]), rows = std::atoi(argv[4]), cols = std::atoi(argv[5]), chnls = std::atoi(argv[6]);
  const int nb = (int)w.strides.size();
  cb_descriptor* desc = nullptr;
  std::vector<cb_ir_block> irb;
  if (w.v2) {  // June2019 MobileNetV2 models: inverted-residual blocks through cb_descriptor_create_v2
    int c_in = 32;
    for (int i = 0; i < nb; ++i) {
      const std::string b = "b" + std::to_string(i) + "_";
      std::vector<int> sh;
      cb_ir_block blk;
      std::memset(&blk, 0, sizeof(blk));
      const std::vector<float>* ew = w.get(b + "expand_w");
      blk.expand_w = ew ? ew->data() : nullptr;
      blk.expand_b = ew ? w.get(b + "expand_b")->data() : nullptr;
      blk.dw_w = w.get(b + "dw_w", &sh)->data();
      blk.dw_b = w.get(b + "dw_b")->data();
      blk.c_exp = sh[2];
      blk.project_w = w.get(b + "project_w", &sh)->data();
      blk.project_b = w.get(b + "project_b")->data();
      blk.c_out = sh[1];
      blk.c_in = c_in;
      blk.stride = w.strides[i];
      blk.residual = w.residual[i];
      c_in = blk.c_out;
      irb.push_back(blk);
    }
    cb_netvlad_v2_weights w2;
    std::vector<int> c1sh, vsh2;
    w2.conv1_w = w.get("conv1_w", &c1sh)->data();
    w2.in_channels = c1sh[2];
    w2.conv1_b = w.get("conv1_b")->data();
    w2.n_blocks = nb;
    w2.blocks = irb.data();
    w2.vlad_w = w.get("vlad_w", &vsh2)->data();
    w2.vlad_d = vsh2[0];
    w2.vlad_k = vsh2[1];
    w2.vlad_b = w.get("vlad_b")->data();
    w2.vlad_c = w.get("vlad_c")->data();
    w2.vlad_ghost = 0;
    if (cb_descriptor_create_v2(&desc, &w2, rows, cols, chnls, 1, 0) != CB_OK) {
      std::fprintf(stderr, "cb_descriptor_create_v2: %s\n", cb_last_error());
      return 1;
    }
  }
  std::vector<const float*> dw_w(nb), dw_b(nb), pw_w(nb), pw_b(nb);
  std::vector<int> cout(nb);
  for (int i = 0; i < nb && !w.v2; ++i) {
    std::vector<int> sh;
    dw_w[i] = w.get("b" + std::to_string(i) + "_dw_w", &sh)->data();
    dw_b[i] = w.get("b" + std::to_string(i) + "_dw_b")->data();
    cout[i] = sh[2];
    const std::vector<float>* pw = w.get("b" + std::to_string(i) + "_pw_w", &sh);
    pw_w[i] = pw ? pw->data() : nullptr;
    pw_b[i] = pw ? w.get("b" + std::to_string(i) + "_pw_b")->data() : nullptr;
    if (pw) cout[i] = sh[1];
  }
  std::vector<int> vsh;
  cb_netvlad_weights nw;
  std::vector<int> c1sh;
  nw.conv1_w = w.get("conv1_w", &c1sh)->data();
  nw.in_channels = c1sh[2];
  nw.n_blocks = nb;
  nw.conv1_b = w.get("conv1_b")->data();
  nw.dw_w = dw_w.data();
  nw.dw_b = dw_b.data();
  nw.pw_w = pw_w.data();
  nw.pw_b = pw_b.data();
  nw.dw_stride = w.strides.data();
  nw.channels_out = cout.data();
  nw.vlad_w = w.get("vlad_w", &vsh)->data();
  nw.vlad_d = vsh[0];
  nw.vlad_k = vsh[1];
  nw.vlad_b = w.get("vlad_b")->data();
  nw.vlad_c = w.get("vlad_c")->data();
  nw.vlad_ghost = 0;  // the shipped models use NetVLADLayer

  if (!w.v2 && cb_descriptor_create(&desc, &nw, rows, cols, chnls, 1, 0) != CB_OK) {
    std::fprintf(stderr, "cb_descriptor_create: %s\n", cb_last_error());
    return 1;
  }
  using namespace cerebro_b200;
  DataManager dm;
  Cerebro cer(desc, rows, cols, chnls);
  if (!cer.ok()) {
    std::fprintf(stderr, "Cerebro init: %s\n", cb_last_error());
    return 1;
  }
  cer.setDataManager(&dm);

  std::ifstream fi(argv[2], std::ios::binary);
  const size_t isz = (size_t)rows * cols * chnls;
  std::string found_json = "[";
  bool first = true;
  for (int i = 0; i < n; ++i) {
    Time t;
    t.nsec = (int64_t)(i + 1) * 100000000LL;  // 10 Hz keyframes
    DataNode* node = new DataNode(t);
    node->left_image.resize(isz);
    fi.read(reinterpret_cast<char*>(node->left_image.data()), (std::streamsize)isz);
    dm.data_map[t] = node;
    if (i % 3 == 2 || i == n - 1) {  // threads wake up: descriptors first, then the search
      cer.descriptor_computer_step();
      const int before = cer.foundLoops_count();
      cer.run_step();
      for (int j = before; j < cer.foundLoops_count(); ++j) {
        auto fl = cer.foundLoops_i(j);
        char buf[128];
        std::snprintf(buf, sizeof(buf), "%s[%lld,%lld,%.17g]", first ? "" : ",",
                      (long long)(std::get<0>(fl).nsec / 100000000LL - 1), (long long)(std::get<1>(fl).nsec / 100000000LL - 1),
                      std::get<2>(fl));
        found_json += buf;
        first = false;
      }
    }
  }
  found_json += "]";
  std::printf("{\"descriptor_size\":%d,\"n_computed\":%d,\"found\":%s", cer.descriptor_size, cer.wholeImageComputedList_size(),
              found_json.c_str());
  if (argc >= 9) {
    const int npts = std::atoi(argv[8]);
    std::vector<double> X((size_t)npts * 3), uv((size_t)npts * 2);
    std::ifstream fp(argv[7], std::ios::binary);
    fp.read(reinterpret_cast<char*>(X.data()), (std::streamsize)(X.size() * 8));
    fp.read(reinterpret_cast<char*>(uv.data()), (std::streamsize)(uv.size() * 8));
    double T[16];
    std::string msg;
    const float conf = cer.verify(X, uv, T, msg);
    std::printf(",\"pnp\":{\"confidence\":%.9g,\"T\":[", conf);
    for (int i = 0; i < 16; ++i) std::printf("%s%.17g", i ? "," : "", T[i]);
    std::printf("]}");
  }
  std::printf("}\n");
  cb_descriptor_destroy(desc);
  return 0;
}
