#include "engine.h"
struct ldn_engine::VaeState {};
namespace ldn {
void vae_finalize(ldn_engine* e, cudaStream_t) { LDN_CHECK(false, "VAE decode not built yet"); }
void vae_decode(ldn_engine*, const float*, float*, int, int, int, cudaStream_t) { LDN_CHECK(false, "VAE decode not built yet"); }
}
