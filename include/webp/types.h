/* webp/types.h -- basic types of the drop-in decode library (libwebp_b200).
 *
 * ABI-compatible with the reference header of the same name (src/webp/types.h:14-72): same macro names, same
 * two allocation entry points. Written for this repository; C99 or C++. */
#ifndef WEBP_WEBP_TYPES_H_
#define WEBP_WEBP_TYPES_H_

#include <stddef.h>
#include <stdint.h>

#ifndef WEBP_INLINE
#define WEBP_INLINE inline
#endif

#ifndef WEBP_EXTERN
#if defined(__GNUC__)
#define WEBP_EXTERN extern __attribute__((visibility("default")))
#else
#define WEBP_EXTERN extern
#endif
#endif

/* Two ABI versions are compatible when their major byte agrees (reference: types.h:53). */
#define WEBP_ABI_IS_INCOMPATIBLE(a, b) (((a) >> 8) != ((b) >> 8))

#ifdef __cplusplus
extern "C" {
#endif

/* malloc/free pair for memory handed out by the WebPDecode*() helpers (reference: types.h:61-66). */
WEBP_EXTERN void* WebPMalloc(size_t size);
WEBP_EXTERN void WebPFree(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* WEBP_WEBP_TYPES_H_ */
