// stubs.cu — placeholders replaced as index types land.
#include "index.h"
namespace b200vs {
#ifndef B200VS_HAVE_HNSW
IndexBase* make_hnsw(b200vs_metric, int, const b200vs_params&) { fail(B200VS_EVECTOR_NOT_SUPPORT, "HNSW not built yet"); }
#endif
}
