#include "fwi_host.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "json_min.hpp"

namespace fwi {

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }
const char *last_error_cstr() { return g_last_error.c_str(); }

static std::string slurp(const std::string &fname, const char *what) {
  std::ifstream f(fname, std::ios::binary);
  if (!f.is_open()) throw Error(FWI_B200_ERR_IO, std::string("cannot open ") + what + " '" + fname + "'");
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

static const JsonValue &need(const JsonValue &o, const char *key, JsonValue::Type t, const char *file) {
  const JsonValue *v = o.find(key);
  if (!v) throw Error(FWI_B200_ERR_JSON, std::string(file) + ": missing key '" + key + "'");
  if (v->type != t) throw Error(FWI_B200_ERR_JSON, std::string(file) + ": key '" + key + "' has the wrong type");
  return *v;
}
static int need_int(const JsonValue &o, const char *key, const char *file) {
  const JsonValue &v = need(o, key, JsonValue::Number, file);
  if (!v.is_int) throw Error(FWI_B200_ERR_JSON, std::string(file) + ": key '" + key + "' must be an integer");
  return static_cast<int>(v.num);
}

Para read_para(const std::string &fname) {
  Para p;
  p.text = slurp(fname, "parameter file");
  JsonValue j;
  try {
    j = JsonParser(p.text).parse();
  } catch (const std::exception &e) {
    throw Error(FWI_B200_ERR_JSON, "parameter file '" + fname + "': " + e.what());
  }
  if (j.type != JsonValue::Object) throw Error(FWI_B200_ERR_JSON, "parameter file is not a JSON object");
  const char *F = "para file";
  p.nz = need_int(j, "nz", F);
  p.nx = need_int(j, "nx", F);
  p.dz = static_cast<float>(need(j, "dz", JsonValue::Number, F).num);
  p.dx = static_cast<float>(need(j, "dx", JsonValue::Number, F).num);
  p.nSteps = need_int(j, "nSteps", F);
  p.nPml = need_int(j, "nPoints_pml", F);
  p.nPad = need_int(j, "nPad", F);
  p.dt = static_cast<float>(need(j, "dt", JsonValue::Number, F).num);
  p.f0 = static_cast<float>(need(j, "f0", JsonValue::Number, F).num);
  p.survey_fname = need(j, "survey_fname", JsonValue::String, F).str;
  p.data_dir_name = need(j, "data_dir_name", JsonValue::String, F).str;
  if (const JsonValue *s = j.find("scratch_dir_name")) {
    if (s->type != JsonValue::String) throw Error(FWI_B200_ERR_JSON, "scratch_dir_name must be a string");
    p.save_scratch = true;
    p.scratch_dir_name = s->str;
  }
  // optional data-conditioning branch (Parameter.cpp:146-177): per-trace windows / weights are built, the
  // band-pass filter and the source-signature update are not -> refused, never ignored
  if (const JsonValue *w = j.find("if_win")) {
    if (w->type != JsonValue::Bool) throw Error(FWI_B200_ERR_JSON, "para file: if_win must be a boolean");
    p.if_win = w->b;
  }
  if (j.has("filter"))
    throw Error(FWI_B200_ERR_UNSUPPORTED, "para file: 'filter' (band-pass) is not supported");
  if (const JsonValue *u = j.find("if_src_update"))
    if (u->type == JsonValue::Bool && u->b)
      throw Error(FWI_B200_ERR_UNSUPPORTED, "para file: if_src_update=true is not supported");
  if (p.nz <= 0 || p.nx <= 0 || p.nSteps < 2 || p.nPml < 0 || p.nPad < 0 || p.nz - p.nPad < 6 || p.nx < 6)
    throw Error(FWI_B200_ERR_GEOM, "para file: inconsistent grid sizes");
  if (!(p.dz > 0) || !(p.dx > 0) || !(p.dt > 0))
    throw Error(FWI_B200_ERR_GEOM, "para file: dz, dx, dt must be positive");
  return p;
}

Survey read_survey(const std::string &fname, int nPml, int group_size, const int *shot_ids, bool if_win) {
  Survey s;
  s.text = slurp(fname, "survey file");
  JsonValue j;
  try {
    j = JsonParser(s.text).parse();
  } catch (const std::exception &e) {
    throw Error(FWI_B200_ERR_JSON, "survey file '" + fname + "': " + e.what());
  }
  if (j.type != JsonValue::Object) throw Error(FWI_B200_ERR_JSON, "survey file is not a JSON object");
  const char *F = "survey file";
  s.nShots = need_int(j, "nShots", F);
  s.shots.resize(group_size);
  for (int i = 0; i < group_size; i++) {
    const std::string key = "shot" + std::to_string(shot_ids[i]);
    const JsonValue *sh = j.find(key);
    if (!sh || sh->type != JsonValue::Object)
      throw Error(FWI_B200_ERR_JSON, "survey file: no entry '" + key + "'");
    Shot &o = s.shots[i];
    o.id = shot_ids[i];
    o.z_src = need_int(*sh, "z_src", F) + nPml;
    o.x_src = need_int(*sh, "x_src", F) + nPml;
    const int nrec = need_int(*sh, "nrec", F);
    const JsonValue &zr = need(*sh, "z_rec", JsonValue::Array, F);
    const JsonValue &xr = need(*sh, "x_rec", JsonValue::Array, F);
    if (nrec < 0 || static_cast<int>(zr.arr.size()) != nrec || static_cast<int>(xr.arr.size()) != nrec)
      throw Error(FWI_B200_ERR_JSON, "survey file: '" + key + "' nrec does not match z_rec / x_rec");
    o.z_rec.resize(nrec);
    o.x_rec.resize(nrec);
    for (int r = 0; r < nrec; r++) {
      if (zr.arr[r].type != JsonValue::Number || xr.arr[r].type != JsonValue::Number)
        throw Error(FWI_B200_ERR_JSON, "survey file: receiver coordinates must be numbers");
      o.z_rec[r] = static_cast<int>(zr.arr[r].num) + nPml;
      o.x_rec[r] = static_cast<int>(xr.arr[r].num) + nPml;
    }
    if (if_win) {  // Src_Rec.cu:157-200: seconds, one value per receiver; doubles narrowed to float
      auto read_f = [&](const char *name, std::vector<float> &dst) {
        const JsonValue &a = need(*sh, name, JsonValue::Array, F);
        if (static_cast<int>(a.arr.size()) != nrec)
          throw Error(FWI_B200_ERR_JSON, "survey file: '" + key + "' " + name + " must have nrec entries");
        dst.resize(nrec);
        for (int r = 0; r < nrec; r++) {
          if (a.arr[r].type != JsonValue::Number) throw Error(FWI_B200_ERR_JSON, std::string("survey file: ") + name + " must be numbers");
          dst[r] = static_cast<float>(a.arr[r].num);
        }
      };
      read_f("win_start", o.win_start);
      read_f("win_end", o.win_end);
      read_f("weights", o.weights);
    }
  }
  return s;
}

// One side of a CPML profile at signed depth `depth` inside the layer.
namespace {
struct Side {
  float damp, K, alpha;
};
}  // namespace

CpmlProfiles cpml_profiles(int N, int nPml, float dh, float f0, float dt) {
  CpmlProfiles p;
  p.K.assign(N, 1.0f);
  p.Kh.assign(N, 1.0f);
  p.a.assign(N, 0.0f);
  p.ah.assign(N, 0.0f);
  p.b.assign(N, 0.0f);
  p.bh.assign(N, 0.0f);
  const double PI = 3.141592653589793238462643383279502884197169;
  const float Rcoef = 0.0008f;
  const float Kmax = 2.0f;
  const float alpha_max = static_cast<float>(2.0 * PI * (static_cast<double>(f0) / 2.0));
  const float npower = 8.0f;
  const float thick = static_cast<float>(nPml) * dh;
  const float cp_ref = 3000.0f;  // the reference pins the PML to 3000 m/s (utilities.cu:259)
  const float d0 = static_cast<float>(static_cast<double>(-(npower + 1.0f) * cp_ref * std::log(Rcoef)) /
                                      (2.0 * static_cast<double>(thick)));
  auto side = [&](float depth, Side &s) {
    const float dn = depth / thick;
    const float p8 = std::pow(dn, npower);
    const float p16 = std::pow(dn, 2.0f * npower);
    s.damp = d0 * (0.25f * dn + 0.75f * p8 + 0.0f * p16);
    s.K = static_cast<float>(1.0 + (static_cast<double>(Kmax) - 1.0) * static_cast<double>(p8));
    s.alpha = static_cast<float>(static_cast<double>(alpha_max) * (1.0 - static_cast<double>(dn)));
  };
  for (int i = 0; i < N; i++) {
    Side full{0.0f, 1.0f, 0.0f}, half{0.0f, 1.0f, 0.0f};
    float depth;
    depth = static_cast<float>(nPml - i) * dh;
    if (depth >= 0.0f) side(depth, full);
    depth = static_cast<float>((static_cast<double>(nPml - i) - 0.5) * static_cast<double>(dh));
    if (depth >= 0.0f) side(depth, half);
    depth = static_cast<float>(nPml - N + i) * dh;
    if (depth >= 0.0f) side(depth, full);
    depth = static_cast<float>((static_cast<double>(nPml - N + i) + 0.5) * static_cast<double>(dh));
    if (depth >= 0.0f) side(depth, half);
    if (full.alpha < 0.0f) full.alpha = 0.0f;
    if (half.alpha < 0.0f) half.alpha = 0.0f;
    p.K[i] = full.K;
    p.Kh[i] = half.K;
    p.b[i] = expf(-(full.damp / full.K + full.alpha) * dt);
    p.bh[i] = expf(-(half.damp / half.K + half.alpha) * dt);
    if (std::fabs(full.damp) > 1.0e-6f)
      p.a[i] = static_cast<float>(static_cast<double>(full.damp) * (static_cast<double>(p.b[i]) - 1.0) /
                                  static_cast<double>(full.K * (full.damp + full.K * full.alpha)));
    if (std::fabs(half.damp) > 1.0e-6f)
      p.ah[i] = static_cast<float>(static_cast<double>(half.damp) * (static_cast<double>(p.bh[i]) - 1.0) /
                                   static_cast<double>(half.K * (half.damp + half.K * half.alpha)));
  }
  return p;
}

bool taper_weights(int nt, float dt, float ratio, std::vector<float> &w2) {
  const double PI = 3.141592653589793238462643383279502884197169;
  w2.assign(nt, 1.0f);
  const float t_end = static_cast<float>(nt) * dt;
  const float ramp = static_cast<float>(nt) * dt * ratio;
  if (2.0 * static_cast<double>(ramp) >= static_cast<double>(t_end)) return false;
  const float t_up = 0.0f + ramp;
  const float t_dn = t_end - ramp;
  for (int k = 0; k < nt; k++) {
    const float t = static_cast<float>(k) * dt;
    float amp;
    if (t >= 0.0f && t < t_up)
      amp = static_cast<float>(std::sin(PI / 2.0 * static_cast<double>(t - 0.0f) / static_cast<double>(t_up - 0.0f)));
    else if (t >= t_up && t < t_dn)
      amp = 1.0f;
    else if (t >= t_dn && t < t_end)
      amp = static_cast<float>(std::cos(PI / 2.0 * static_cast<double>(t - t_dn) / static_cast<double>(t_end - t_dn)));
    else
      amp = 0.0f;
    w2[k] = amp * amp;
  }
  return true;
}

float courant_number(float cp_max, float dt, float dz, float dx) {
  const float dh = (dz < dx) ? dz : dx;
  return static_cast<float>(static_cast<double>(cp_max * dt * sqrtf(2.0f)) * (1.0 / 24.0 + 9.0 / 8.0) /
                            static_cast<double>(dh));
}

void read_f32(const std::string &fname, float *dst, size_t n) {
  FILE *fp = std::fopen(fname.c_str(), "rb");
  if (!fp) throw Error(FWI_B200_ERR_IO, "cannot read '" + fname + "'");
  std::memset(dst, 0, n * sizeof(float));
  size_t got = std::fread(dst, sizeof(float), n, fp);
  (void)got;  // a short file leaves zeros behind, like the reference (utilities.cu:10-19)
  std::fclose(fp);
}

void write_f32(const std::string &fname, const float *src, size_t n) {
  FILE *fp = std::fopen(fname.c_str(), "wb");
  if (!fp) throw Error(FWI_B200_ERR_IO, "cannot write '" + fname + "'");
  size_t put = std::fwrite(src, sizeof(float), n, fp);
  std::fclose(fp);
  if (put != n) throw Error(FWI_B200_ERR_IO, "short write to '" + fname + "'");
}

}  // namespace fwi
