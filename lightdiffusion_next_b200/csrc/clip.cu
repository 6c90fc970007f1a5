#include "engine.h"
struct ldn_engine::ClipState {};
namespace ldn {
void clip_finalize(ldn_engine* e, cudaStream_t) { LDN_CHECK(false, "CLIP encode not built yet"); }
void clip_encode(ldn_engine*, const int64_t*, int, float*, float*, cudaStream_t) { LDN_CHECK(false, "CLIP encode not built yet"); }
}
