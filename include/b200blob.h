/*
 * b200blob.h -- host side of the non-Python SDKs' blob upload, above the C ABI of b200hash.h.
 *
 * The Go and JS SDKs of the reference hash a whole payload once with their standard libraries, base64 the
 * digests, call BlobCreate and PUT the bytes (go/blob.go:50-95 `blobUpload`, js/src/blob.ts:31-69), from
 * `createInput` when the CBOR-encoded arguments exceed 2 MiB (go/function.go:193-201).  Neither toolchain exists
 * in this image, so the host logic is mirrored here in C++ behind a C interface a cgo / N-API binding can call
 * directly: same steps, same retry policy (3 attempts, 300 ms doubling), same error texts.  The RPC and the HTTP
 * PUT stay with the caller (callbacks): they are the SDK's transport, not part of the hash path.
 *
 * Functions return 0 on success or a negative B200BLOB_E_* code; the message goes to err_out (B200BLOB_ERR_MAX).
 */
#ifndef B200BLOB_H
#define B200BLOB_H

#include "b200hash.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B200BLOB_ID_MAX 128
#define B200BLOB_URL_MAX 4096
#define B200BLOB_ERR_MAX 256
#define B200BLOB_MD5_B64_LEN 25    /* 24 characters + NUL */
#define B200BLOB_SHA256_B64_LEN 45 /* 44 characters + NUL */

#define B200BLOB_E_HASH (-1)      /* the GPU hash call failed (message from b200h_last_error) */
#define B200BLOB_E_CREATE (-2)    /* "failed to create blob: ..." (go/blob.go:62) */
#define B200BLOB_E_MULTIPART (-3) /* multipart response: unsupported by these SDKs (go/blob.go:66-67, blob.ts:42-45) */
#define B200BLOB_E_NO_URL (-4)    /* "missing upload URL in BlobCreate response" (go/blob.go:92-93) */
#define B200BLOB_E_PUT (-5)       /* "failed blob upload: <status>" after the retries (go/blob.go:85-87) */
#define B200BLOB_E_INVALID (-6)

enum { B200BLOB_UPLOAD_NONE = 0, B200BLOB_UPLOAD_URL = 1, B200BLOB_UPLOAD_MULTIPART = 2 };

/* What the caller's BlobCreate RPC returned (modal_proto/api.proto BlobCreateResponse). */
typedef struct {
    int upload_type; /* B200BLOB_UPLOAD_*: which member of upload_type_oneof is set */
    char blob_id[B200BLOB_ID_MAX];
    char upload_url[B200BLOB_URL_MAX];
} b200blob_create_response;

typedef struct {
    /* BlobCreate(content_md5, content_sha256_base64, content_length): 0 = ok, nonzero = RPC error (text in err). */
    int (*blob_create)(void* user, const char* content_md5, const char* content_sha256_base64,
                       int64_t content_length, b200blob_create_response* out, char err[B200BLOB_ERR_MAX]);
    /* PUT `data` to `url` with Content-Type: application/octet-stream and Content-MD5: content_md5.
     * Returns the HTTP status, or a negative number for a transport error. */
    int (*http_put)(void* user, const char* url, const uint8_t* data, uint64_t len, const char* content_md5);
    /* Back-off between attempts; NULL = the library sleeps itself. */
    void (*sleep_ms)(void* user, unsigned ms);
    void* user;
} b200blob_transport;

/* createInput's gate (go/function.go:193): payloads strictly larger than 2 MiB go to blob storage. */
int b200blob_should_upload(uint64_t nbytes);

/* MD5 and SHA-256 of n payloads in ONE GPU batch, as standard base64 with padding
 * (md5.Sum / sha256.Sum256 + base64.StdEncoding of go/blob.go:51-54, createHash().digest("base64") of blob.ts:35-36).
 * md5_b64_out: n * B200BLOB_MD5_B64_LEN chars, sha256_b64_out: n * B200BLOB_SHA256_B64_LEN chars. */
int b200blob_hashes_many(b200h_ctx* ctx, const uint8_t* const* data, const uint64_t* len, uint64_t n,
                         char* md5_b64_out, char* sha256_b64_out, char err_out[B200BLOB_ERR_MAX]);

/* blobUpload (go/blob.go:50-95): hash on the GPU, BlobCreate, PUT with up to 3 attempts.  blob_id_out receives
 * the blob id on success. */
int b200blob_upload(b200h_ctx* ctx, const b200blob_transport* transport, const uint8_t* data, uint64_t len,
                    char blob_id_out[B200BLOB_ID_MAX], char err_out[B200BLOB_ERR_MAX]);

/* Many payloads (a map's worth of inputs): all digests from one GPU batch, then create + PUT per payload in order.
 * status_out[i] = 0 or the B200BLOB_E_* code of payload i; blob_ids_out = n * B200BLOB_ID_MAX chars.  Returns 0 when
 * every payload succeeded, else the first failure's code (its message in err_out). */
int b200blob_upload_many(b200h_ctx* ctx, const b200blob_transport* transport, const uint8_t* const* data,
                         const uint64_t* len, uint64_t n, char* blob_ids_out, int* status_out,
                         char err_out[B200BLOB_ERR_MAX]);

#ifdef __cplusplus
}
#endif
#endif /* B200BLOB_H */
