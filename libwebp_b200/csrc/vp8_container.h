/* vp8_container.h -- host-side container / frame-tag parsing (see vp8_container.c). */
#ifndef LIBWEBP_B200_VP8_CONTAINER_H_
#define LIBWEBP_B200_VP8_CONTAINER_H_

#include "webp/decode.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Vp8Container {
  int width, height;
  int has_alpha, has_animation, format;   /* WebPBitstreamFeatures */
  int found_vp8x, is_lossless, has_alph_chunk;
  int complete;              /* the walk reached the image bitstream */
  size_t frame_offset;       /* offset of the VP8 frame tag (or VP8L signature) in the file */
  size_t frame_size;         /* bytes from there to the end of the input (reference: io.data_size) */
  size_t chunk_size;         /* declared payload size */
  size_t alpha_offset, alpha_size;
  uint32_t part0_size;       /* first-partition length from the frame tag */
} Vp8Container;

/* Walks the container. have_all_data = 1 for a decode (WebPParseHeaders), 0 for a feature probe. */
int vp8b_parse_container(const uint8_t* data, size_t size, int have_all_data, Vp8Container* out);

/* WebPGetFeatures semantics (src/dec/webp_dec.c:684-704). */
VP8StatusCode vp8b_get_features(const uint8_t* data, size_t size, WebPBitstreamFeatures* f);

/* Number of token partitions (1, 2, 4 or 8) announced in the first partition. */
int vp8b_prescan_partitions(const uint8_t* part0, size_t part0_size);

/* 1 once the last token partition has begun inside the `rest` bytes at hand after the frame's 10-byte header. */
int vp8b_last_partition_begun(const uint8_t* part0, size_t part0_size, size_t rest, int num_parts);

/* 1 when the first partition or a token partition starts with byte 0xFF (`rest` = frame bytes after the 10-byte header). */
int vp8b_partition_starts_with_ff(const uint8_t* part0, size_t part0_size, size_t rest, int num_parts);

#ifdef __cplusplus
}
#endif
#endif /* LIBWEBP_B200_VP8_CONTAINER_H_ */
