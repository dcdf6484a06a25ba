// Second build of the fused highlight kernel (highlight_fused.cu): 256-thread CTAs with 48 KB of dynamic shared memory,
// four to an SM, for small frames.  Same source, same results; see the variant note at the top of highlight_fused.cu.
#define CVVP_HLF_SMALL 1
#include "highlight_fused.cu"
