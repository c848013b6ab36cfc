// oracle/ref_glue_cuda.cpp -- TEST INFRASTRUCTURE ONLY.  The same translation unit as ref_glue.cpp with the drop-in
// adapter compiled in: the reference's GCRANSAC<> driver, unmodified, with gcr::CudaMSACScoringFunction as its
// _ScoringFunction template argument.  Links libgcr_b200.so.  Built by oracle/build_ref.sh into
// oracle/_ref/libgcr_ref_cuda.so (only where /root/reference exists).
#define GCR_REF_WITH_CUDA 1
#include "ref_glue.cpp"
