// TEST INFRASTRUCTURE (oracle build shim) — not product code.
// Stand-in for lodepng so the unmodified reference src/user/user_objects.cc
// compiles.  PNG decoding is only used for textures / height fields, never on
// the physics path; decode always reports an error.
#ifndef ORACLE_SHIM_LODEPNG_H_
#define ORACLE_SHIM_LODEPNG_H_
#include <cstddef>
typedef enum LodePNGColorType { LCT_GREY = 0, LCT_RGB = 2, LCT_PALETTE = 3, LCT_GREY_ALPHA = 4, LCT_RGBA = 6 } LodePNGColorType;
typedef struct LodePNGColorMode { LodePNGColorType colortype; unsigned bitdepth; } LodePNGColorMode;
namespace lodepng {
struct Info { unsigned srgb_defined = 0; };
struct State { LodePNGColorMode info_raw{LCT_RGBA, 8}; Info info_png; };
}
inline unsigned lodepng_decode(unsigned char** out, unsigned* w, unsigned* h, lodepng::State*,
                               const unsigned char*, size_t) {
  *out = nullptr; *w = 0; *h = 0; return 1;
}
inline const char* lodepng_error_text(unsigned) { return "PNG decoding is not available in the oracle build"; }
inline size_t lodepng_get_raw_size(unsigned w, unsigned h, const LodePNGColorMode* m) {
  size_t ch = m->colortype == LCT_GREY ? 1 : m->colortype == LCT_RGB ? 3 : m->colortype == LCT_GREY_ALPHA ? 2 : 4;
  return (size_t)w * h * ch;
}
#endif
