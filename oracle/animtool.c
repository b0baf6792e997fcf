// TEST INFRASTRUCTURE ONLY. Animated WebP through the reference's own caller of the decode path:
// WebPAnimDecoder (src/demux/anim_decode.c:376 WebPAnimDecoderGetNext) demuxes a file, calls WebPDecode once per frame into a
// sub-rectangle of its canvas (external memory, canvas stride) and blends on the host. This file is compiled twice by
// oracle/Makefile against the UNMODIFIED reference demuxer:
//   * into oracle/_ref/libwebp_ref.so with ANIM_PREFIX=reft_   -> every WebPDecode is the reference's (the oracle);
//   * into oracle/_ref/libanim_b200.so with ANIM_PREFIX=b200_  -> linked against libwebpdecoder_b200.so, so the very same
//     caller code drives the CUDA decoder (the demux / animation caller above the drop-in boundary).
// The encoder half (reference WebPAnimEncoder, src/mux/anim_encode.c) only exists in the reft_ build.
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "src/webp/decode.h"
#include "src/webp/demux.h"

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(ANIM_PREFIX, name)

// Which shared object the WebPDecode this build calls lives in (the test asserts it is the one it means to exercise).
const char* FN(anim_decoder_library)(void) {
  Dl_info info;
  if (!dladdr((void*)&WebPDecode, &info) || info.dli_fname == NULL) return "";
  return info.dli_fname;
}

// info3 = {canvas width, canvas height, frame count}. Returns 1 on success.
int FN(anim_info)(const uint8_t* data, size_t size, int* info3) {
  WebPData d = { data, size };
  WebPAnimDecoderOptions o;
  WebPAnimInfo ai;
  WebPAnimDecoder* dec;
  if (!WebPAnimDecoderOptionsInit(&o)) return 0;
  dec = WebPAnimDecoderNew(&d, &o);
  if (dec == NULL) return 0;
  if (!WebPAnimDecoderGetInfo(dec, &ai)) { WebPAnimDecoderDelete(dec); return 0; }
  info3[0] = (int)ai.canvas_width; info3[1] = (int)ai.canvas_height; info3[2] = (int)ai.frame_count;
  WebPAnimDecoderDelete(dec);
  return 1;
}

// Every reconstructed canvas (4 * w * h bytes each) one after the other into `out`, time stamps into `ts`.
// Returns the number of frames delivered, or -1 - (frames delivered) when WebPAnimDecoderGetNext failed.
int FN(anim_decode_all)(const uint8_t* data, size_t size, int csp, uint8_t* out, size_t out_size, int* ts) {
  WebPData d = { data, size };
  WebPAnimDecoderOptions o;
  WebPAnimInfo ai;
  WebPAnimDecoder* dec;
  int n = 0;
  if (!WebPAnimDecoderOptionsInit(&o)) return -1;
  o.color_mode = (WEBP_CSP_MODE)csp;
  dec = WebPAnimDecoderNew(&d, &o);
  if (dec == NULL) return -1;
  if (!WebPAnimDecoderGetInfo(dec, &ai)) { WebPAnimDecoderDelete(dec); return -1; }
  {
    const size_t frame_bytes = (size_t)4 * ai.canvas_width * ai.canvas_height;
    while (WebPAnimDecoderHasMoreFrames(dec)) {
      uint8_t* canvas;
      int t;
      if (!WebPAnimDecoderGetNext(dec, &canvas, &t)) { n = -1 - n; break; }
      if ((size_t)(n + 1) * frame_bytes > out_size) { n = -1 - n; break; }
      memcpy(out + (size_t)n * frame_bytes, canvas, frame_bytes);
      if (ts != NULL) ts[n] = t;
      ++n;
    }
  }
  WebPAnimDecoderDelete(dec);
  return n;
}

#ifdef ANIM_WITH_ENCODER
#include "src/webp/encode.h"
#include "src/webp/mux.h"
// n RGBA frames (w x h, tightly packed, one after the other) -> animated WebP, key frames every `kmax` frames at most so that
// sub-rectangle frames with blending and ALPH chunks occur. lossless: 0 = every frame lossy at `quality`, 1 = every frame
// lossless, 2 = the encoder picks per frame (allow_mixed).
size_t FN(anim_encode)(const uint8_t* frames, int n, int w, int h, float quality, int method, int kmin, int kmax,
                       int minimize_size, int lossless, uint8_t** out) {
  WebPAnimEncoderOptions eo;
  WebPAnimEncoder* enc;
  WebPData wd;
  size_t size = 0;
  int i, ok = 1;
  *out = NULL;
  if (!WebPAnimEncoderOptionsInit(&eo)) return 0;
  eo.kmin = kmin; eo.kmax = kmax; eo.minimize_size = minimize_size; eo.allow_mixed = (lossless == 2);
  enc = WebPAnimEncoderNew(w, h, &eo);
  if (enc == NULL) return 0;
  for (i = 0; ok && i < n; ++i) {
    WebPConfig config;
    WebPPicture pic;
    if (!WebPConfigInit(&config) || !WebPPictureInit(&pic)) { ok = 0; break; }
    config.lossless = (lossless == 1); config.quality = quality; config.method = method;
    pic.use_argb = 1; pic.width = w; pic.height = h;
    if (!WebPPictureImportRGBA(&pic, frames + (size_t)i * 4 * w * h, 4 * w)) { ok = 0; break; }
    ok = WebPAnimEncoderAdd(enc, &pic, 40 * i, &config);
    WebPPictureFree(&pic);
  }
  if (ok) ok = WebPAnimEncoderAdd(enc, NULL, 40 * n, NULL);
  WebPDataInit(&wd);
  if (ok && WebPAnimEncoderAssemble(enc, &wd)) {
    *out = (uint8_t*)malloc(wd.size);
    if (*out != NULL) { memcpy(*out, wd.bytes, wd.size); size = wd.size; }
  }
  WebPDataClear(&wd);
  WebPAnimEncoderDelete(enc);
  return size;
}
#endif
