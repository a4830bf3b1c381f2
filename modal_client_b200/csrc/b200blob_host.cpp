// b200blob_host.cpp -- C++ mirror of the Go / JS SDKs' blobUpload host logic above the C ABI (include/b200blob.h).
// Follows go/blob.go:18-95 step by step; hashing is the only part that differs: one GPU batch
// (b200h_hash_batch_host with absolute addresses) instead of md5.Sum + sha256.Sum256 per payload.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/b200blob.h"

namespace {

constexpr int kUploadRetryAttempts = 3;          // go/blob.go:19
constexpr unsigned kUploadRetryDelayMs = 300;    // go/blob.go:20
constexpr uint64_t kMaxObjectSizeBytes = 2 * 1024 * 1024;  // go/function.go:27, js/src/function.ts

void set_err(char* err, const char* fmt, const char* a) {
    if (err) snprintf(err, B200BLOB_ERR_MAX, fmt, a);
}

// base64.StdEncoding: alphabet A-Za-z0-9+/ with '=' padding
void b64(const uint8_t* in, size_t n, char* out) {
    static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    size_t o = 0;
    for (size_t i = 0; i < n; i += 3) {
        const uint32_t b0 = in[i], b1 = i + 1 < n ? in[i + 1] : 0, b2 = i + 2 < n ? in[i + 2] : 0;
        const uint32_t v = (b0 << 16) | (b1 << 8) | b2;
        out[o++] = T[(v >> 18) & 63];
        out[o++] = T[(v >> 12) & 63];
        out[o++] = i + 1 < n ? T[(v >> 6) & 63] : '=';
        out[o++] = i + 2 < n ? T[v & 63] : '=';
    }
    out[o] = 0;
}

int put_with_retries(const b200blob_transport* t, const b200blob_create_response& resp, const uint8_t* data,
                     uint64_t len, const char* md5_b64, char* err) {
    // retryHTTPRequest (go/blob.go:26-48): attempts with exponential back-off, last error wins
    unsigned delay = kUploadRetryDelayMs;
    int status = 0;
    for (int attempt = 0; attempt < kUploadRetryAttempts; ++attempt) {
        status = t->http_put(t->user, resp.upload_url, data, len, md5_b64);
        if (status >= 200 && status < 300) return 0;
        if (attempt < kUploadRetryAttempts - 1) {
            if (t->sleep_ms) t->sleep_ms(t->user, delay);
            else std::this_thread::sleep_for(std::chrono::milliseconds(delay));
            delay *= 2;
        }
    }
    if (err) {
        if (status < 0) snprintf(err, B200BLOB_ERR_MAX, "failed to upload blob: transport error %d", status);
        else snprintf(err, B200BLOB_ERR_MAX, "failed blob upload: %d", status);
    }
    return B200BLOB_E_PUT;
}

int create_and_put(const b200blob_transport* t, const uint8_t* data, uint64_t len, const char* md5_b64,
                   const char* sha_b64, char* blob_id_out, char* err) {
    b200blob_create_response resp;
    memset(&resp, 0, sizeof resp);
    char rpc_err[B200BLOB_ERR_MAX] = "";
    if (t->blob_create(t->user, md5_b64, sha_b64, (int64_t)len, &resp, rpc_err) != 0) {
        set_err(err, "failed to create blob: %s", rpc_err);
        return B200BLOB_E_CREATE;
    }
    switch (resp.upload_type) {
        case B200BLOB_UPLOAD_MULTIPART:
            set_err(err, "%s", "Function input size exceeds multipart upload threshold, unsupported by this SDK version");
            return B200BLOB_E_MULTIPART;
        case B200BLOB_UPLOAD_URL: {
            if (int rc = put_with_retries(t, resp, data, len, md5_b64, err)) return rc;
            resp.blob_id[B200BLOB_ID_MAX - 1] = 0;
            memcpy(blob_id_out, resp.blob_id, B200BLOB_ID_MAX);
            return 0;
        }
        default:
            set_err(err, "%s", "missing upload URL in BlobCreate response");
            return B200BLOB_E_NO_URL;
    }
}

}  // namespace

extern "C" {

int b200blob_should_upload(uint64_t nbytes) { return nbytes > kMaxObjectSizeBytes; }

int b200blob_hashes_many(b200h_ctx* ctx, const uint8_t* const* data, const uint64_t* len, uint64_t n,
                         char* md5_b64_out, char* sha256_b64_out, char err_out[B200BLOB_ERR_MAX]) {
    if (!ctx || (n && (!data || !len || !md5_b64_out || !sha256_b64_out))) {
        set_err(err_out, "%s", "null argument");
        return B200BLOB_E_INVALID;
    }
    if (n == 0) return 0;
    std::vector<uint64_t> addr(n);
    for (uint64_t i = 0; i < n; ++i) addr[i] = (uint64_t)(uintptr_t)data[i];  // absolute addresses, base = NULL
    std::vector<uint8_t> sha(n * 32), md5(n * 16);
    const int rc = b200h_hash_batch_host(ctx, nullptr, addr.data(), len, n, B200H_SHA256 | B200H_MD5, sha.data(),
                                         md5.data(), nullptr);
    if (rc != 0) {
        set_err(err_out, "GPU hash failed: %s", b200h_last_error(ctx));
        return B200BLOB_E_HASH;
    }
    for (uint64_t i = 0; i < n; ++i) {
        b64(&md5[i * 16], 16, md5_b64_out + i * B200BLOB_MD5_B64_LEN);
        b64(&sha[i * 32], 32, sha256_b64_out + i * B200BLOB_SHA256_B64_LEN);
    }
    return 0;
}

int b200blob_upload(b200h_ctx* ctx, const b200blob_transport* transport, const uint8_t* data, uint64_t len,
                    char blob_id_out[B200BLOB_ID_MAX], char err_out[B200BLOB_ERR_MAX]) {
    if (!transport || !transport->blob_create || !transport->http_put || !blob_id_out) {
        set_err(err_out, "%s", "null argument");
        return B200BLOB_E_INVALID;
    }
    char md5_b64[B200BLOB_MD5_B64_LEN], sha_b64[B200BLOB_SHA256_B64_LEN];
    if (int rc = b200blob_hashes_many(ctx, &data, &len, 1, md5_b64, sha_b64, err_out)) return rc;
    return create_and_put(transport, data, len, md5_b64, sha_b64, blob_id_out, err_out);
}

int b200blob_upload_many(b200h_ctx* ctx, const b200blob_transport* transport, const uint8_t* const* data,
                         const uint64_t* len, uint64_t n, char* blob_ids_out, int* status_out,
                         char err_out[B200BLOB_ERR_MAX]) {
    if (!transport || !transport->blob_create || !transport->http_put || (n && (!blob_ids_out || !status_out))) {
        set_err(err_out, "%s", "null argument");
        return B200BLOB_E_INVALID;
    }
    std::vector<char> md5_b64(n * B200BLOB_MD5_B64_LEN + 1), sha_b64(n * B200BLOB_SHA256_B64_LEN + 1);
    if (int rc = b200blob_hashes_many(ctx, data, len, n, md5_b64.data(), sha_b64.data(), err_out)) return rc;
    int first_rc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        char e[B200BLOB_ERR_MAX] = "";
        char* id = blob_ids_out + i * B200BLOB_ID_MAX;
        id[0] = 0;
        status_out[i] = create_and_put(transport, data[i], len[i], &md5_b64[i * B200BLOB_MD5_B64_LEN],
                                       &sha_b64[i * B200BLOB_SHA256_B64_LEN], id, e);
        if (status_out[i] != 0 && first_rc == 0) {
            first_rc = status_out[i];
            if (err_out) memcpy(err_out, e, B200BLOB_ERR_MAX);
        }
    }
    return first_rc;
}

}  // extern "C"
