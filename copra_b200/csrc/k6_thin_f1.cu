// k6_thin_f1.cu -- gi_thin_kernel<512, {1,2}, 1> (see gi_thin_launch.cuh)
#include "gi_thin_launch.cuh"

namespace cb {
GT_DEFINE_FORM_LAUNCH(1)
}
