// placeholder until the fused scan lands
#include "internal.h"
namespace vmb {
bool scan_fast_supported(const FastScanArgs&) { return false; }
int scan_fast(const FastScanArgs&, cudaStream_t) { VMB_UNSUPPORTED("scan_fast: not built"); }
}
