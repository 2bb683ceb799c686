// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" doorway into the UNMODIFIED reference CUDA library (built by
// oracle/build_ref.sh from the sources where they lie under /root/reference;
// no reference source is copied into this repo).  The reference entry point is
// the C++ symbol `cufd(..., const std::string para_fname)`
// (deps/CustomOps/FWI/Src/libCUFD.cu:34-38, declared in
// deps/CustomOps/FWI/FwiOp.h:9-12); a C caller (ctypes) cannot build a
// std::string, hence this one-function shim.
#include <string>

void cufd(double *misfit, double *grad_Lambda, double *grad_Mu,
          double *grad_Den, double *grad_stf, const double *Lambda,
          const double *Mu, const double *Den, const double *stf, int calc_id,
          const int gpu_id, int group_size, const int *shot_ids,
          const std::string para_fname);

extern "C" int ref_cufd(double *misfit, double *grad_Lambda, double *grad_Mu,
                        double *grad_Den, double *grad_stf,
                        const double *Lambda, const double *Mu,
                        const double *Den, const double *stf, int calc_id,
                        int gpu_id, int group_size, const int *shot_ids,
                        const char *para_fname) {
  cufd(misfit, grad_Lambda, grad_Mu, grad_Den, grad_stf, Lambda, Mu, Den, stf,
       calc_id, gpu_id, group_size, shot_ids, std::string(para_fname));
  return 0;
}
