// Host-side pieces of the FWI path: parameter / survey files, CPML profiles,
// time tapers, Shot<id>.bin I/O, error plumbing.  C++17, no CUDA in here.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/fwi_b200.h"

namespace fwi {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);
const char *last_error_cstr();

// ---- para_file.json (reference: Parameter.cpp:16-178) -----------------------
struct Para {
  int nz = 0, nx = 0, nSteps = 0, nPml = 0, nPad = 0;
  float dz = 0, dx = 0, dt = 0, f0 = 0;
  std::string survey_fname, data_dir_name, scratch_dir_name;
  bool save_scratch = false;
  bool if_win = false;  // per-trace time windows + trace weights from the survey file (Parameter.cpp:146-150)
  std::string text;  // raw file content (plan-cache key)
};
Para read_para(const std::string &fname);

// ---- survey_file.json (reference: Src_Rec.cu:19-115) ------------------------
struct Shot {
  int id = 0;            // global shot id ("shot<id>")
  int z_src = 0, x_src = 0;  // padded, 0-based (json value + nPml)
  std::vector<int> z_rec, x_rec;  // padded, 0-based
  std::vector<float> win_start, win_end, weights;  // per receiver, only with if_win (Src_Rec.cu:157-200)
};
struct Survey {
  int nShots = 0;
  std::vector<Shot> shots;  // the shots of the group, in group order
  std::string text;
};
Survey read_survey(const std::string &fname, int nPml, int group_size, const int *shot_ids, bool if_win = false);

// ---- CPML profiles (reference: utilities.cu:242-358, Cpml.cu:46-52) ---------
struct CpmlProfiles {
  std::vector<float> K, a, b, Kh, ah, bh;
};
CpmlProfiles cpml_profiles(int N, int nPml, float dh, float f0, float dt);

// ---- time taper (reference: utilities.cu:707-747, ratio 0.005 / 0.001) ------
// w2[t] = window_amp(t)^2 as float; returns false on the reference's "Window error 2"
// (the data are then left untouched).
bool taper_weights(int nt, float dt, float ratio, std::vector<float> &w2);

// Courant number from max cp (reference: utilities.cu:225-240)
float courant_number(float cp_max, float dt, float dz, float dx);

// ---- raw float32 files (reference: utilities.cu:10-30) ----------------------
void read_f32(const std::string &fname, float *dst, size_t n);
void write_f32(const std::string &fname, const float *src, size_t n);

}  // namespace fwi
