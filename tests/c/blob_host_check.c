/* tests/c/blob_host_check.c -- a plain C99 host (what a cgo binding compiles to) driving include/b200blob.h and
 * include/b200hash.h through fake transport callbacks.  Built and run by tests/test_c_host.py.
 *   blob_host_check nogpu          -> checks that need no device (size gate, loud failure without a GPU)
 *   blob_host_check run            -> uploads synthetic payloads; prints one line per fact for the Python side,
 *                                     which compares the digests with hashlib on the same bytes. */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200blob.h"

/* byte i of payload `seed`: what tests/test_c_host.py regenerates with numpy */
static void fill(uint8_t* p, uint64_t n, uint32_t seed) {
    for (uint64_t i = 0; i < n; ++i) p[i] = (uint8_t)((((uint32_t)i + seed) * 2654435761u) >> 24);
}

typedef struct {
    int creates, puts, sleeps;
    unsigned slept_ms;
    int fail_puts_remaining;   /* respond 500 this many times before 200 */
    int64_t multipart_above;   /* BlobCreate answers "multipart" above this length */
    int rpc_fails, no_url;
    const uint8_t* last_data;
    uint64_t last_len;
    char last_md5[B200BLOB_MD5_B64_LEN], last_sha[B200BLOB_SHA256_B64_LEN], last_put_md5[B200BLOB_MD5_B64_LEN];
} fake;

static int fake_create(void* user, const char* md5, const char* sha, int64_t len, b200blob_create_response* out,
                       char err[B200BLOB_ERR_MAX]) {
    fake* f = (fake*)user;
    f->creates++;
    snprintf(f->last_md5, sizeof f->last_md5, "%s", md5);
    snprintf(f->last_sha, sizeof f->last_sha, "%s", sha);
    if (f->rpc_fails) {
        snprintf(err, B200BLOB_ERR_MAX, "rpc unavailable");
        return 1;
    }
    snprintf(out->blob_id, sizeof out->blob_id, "bl-%d", f->creates);
    if (f->no_url) {
        out->upload_type = B200BLOB_UPLOAD_NONE;
    } else if (len > f->multipart_above) {
        out->upload_type = B200BLOB_UPLOAD_MULTIPART;
    } else {
        out->upload_type = B200BLOB_UPLOAD_URL;
        snprintf(out->upload_url, sizeof out->upload_url, "http://fake/upload?blob_id=bl-%d", f->creates);
    }
    return 0;
}

static int fake_put(void* user, const char* url, const uint8_t* data, uint64_t len, const char* md5) {
    fake* f = (fake*)user;
    f->puts++;
    f->last_data = data;
    f->last_len = len;
    snprintf(f->last_put_md5, sizeof f->last_put_md5, "%s", md5);
    if (strncmp(url, "http://fake/upload?blob_id=bl-", 30) != 0) return -1;
    if (f->fail_puts_remaining > 0) {
        f->fail_puts_remaining--;
        return 500;
    }
    return 200;
}

static void fake_sleep(void* user, unsigned ms) {
    fake* f = (fake*)user;
    f->sleeps++;
    f->slept_ms += ms;
}

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            printf("CHECK FAILED line %d: %s\n", __LINE__, #cond);         \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "nogpu";
    CHECK(!b200blob_should_upload(2 * 1024 * 1024));     /* strict > (go/function.go:193) */
    CHECK(b200blob_should_upload(2 * 1024 * 1024 + 1));
    printf("version %s\n", b200h_version());

    b200h_ctx* ctx = NULL;
    int rc = b200h_create(0, 8u << 20, 64u << 20, &ctx);
    if (strcmp(mode, "nogpu") == 0) {
        if (b200h_device_count() <= 0) {
            CHECK(rc == B200H_E_CUDA && ctx == NULL);
            CHECK(strlen(b200h_last_error(NULL)) > 0);
            printf("create failed as it must: %s\n", b200h_last_error(NULL));
        } else if (ctx) {
            b200h_destroy(ctx);
        }
        printf("OK nogpu\n");
        return 0;
    }
    if (rc != 0) {
        printf("b200h_create failed: %s\n", b200h_last_error(NULL));
        return 2;
    }

    fake f;
    memset(&f, 0, sizeof f);
    f.multipart_above = 1 << 20;
    b200blob_transport t = {fake_create, fake_put, fake_sleep, &f};
    char id[B200BLOB_ID_MAX], err[B200BLOB_ERR_MAX];

    /* 1. single uploads of assorted sizes; digests printed for the hashlib comparison */
    const uint64_t sizes[] = {0, 1, 55, 56, 64, 1000, 65537, 1 << 20};
    for (size_t k = 0; k < sizeof sizes / sizeof sizes[0]; ++k) {
        uint8_t* p = (uint8_t*)malloc(sizes[k] ? sizes[k] : 1);
        fill(p, sizes[k], (uint32_t)(100 + k));
        err[0] = 0;
        rc = b200blob_upload(ctx, &t, p, sizes[k], id, err);
        CHECK(rc == 0);
        CHECK(f.last_data == p && f.last_len == sizes[k]);          /* the PUT carries the caller's bytes */
        CHECK(strcmp(f.last_md5, f.last_put_md5) == 0);              /* Content-MD5 == BlobCreate content_md5 */
        printf("hashes %u %" PRIu64 " %s %s %s\n", (unsigned)(100 + k), sizes[k], f.last_md5, f.last_sha, id);
        free(p);
    }
    /* 2. retry policy: two 500s then success = 3 attempts, 300 ms + 600 ms of back-off */
    {
        uint8_t p[100];
        fill(p, sizeof p, 7);
        const int puts0 = f.puts, sleeps0 = f.sleeps;
        const unsigned ms0 = f.slept_ms;
        f.fail_puts_remaining = 2;
        CHECK(b200blob_upload(ctx, &t, p, sizeof p, id, err) == 0);
        CHECK(f.puts - puts0 == 3 && f.sleeps - sleeps0 == 2 && f.slept_ms - ms0 == 900);
        f.fail_puts_remaining = 3; /* all three attempts fail */
        CHECK(b200blob_upload(ctx, &t, p, sizeof p, id, err) == B200BLOB_E_PUT);
        printf("put_error %s\n", err);
        f.fail_puts_remaining = 0;
    }
    /* 3. multipart / missing URL / RPC error texts of go/blob.go */
    {
        uint8_t* p = (uint8_t*)malloc((1 << 20) + 1);
        fill(p, (1 << 20) + 1, 9);
        CHECK(b200blob_upload(ctx, &t, p, (1 << 20) + 1, id, err) == B200BLOB_E_MULTIPART);
        printf("multipart_error %s\n", err);
        f.no_url = 1;
        CHECK(b200blob_upload(ctx, &t, p, 10, id, err) == B200BLOB_E_NO_URL);
        printf("no_url_error %s\n", err);
        f.no_url = 0;
        f.rpc_fails = 1;
        CHECK(b200blob_upload(ctx, &t, p, 10, id, err) == B200BLOB_E_CREATE);
        printf("create_error %s\n", err);
        f.rpc_fails = 0;
        free(p);
    }
    /* 4. a map's worth of inputs: one GPU batch, ids in order, one oversized payload fails alone */
    {
        enum { N = 300 };
        const uint8_t* ptrs[N];
        uint64_t lens[N];
        static char ids[N * B200BLOB_ID_MAX];
        int status[N];
        for (int i = 0; i < N; ++i) {
            lens[i] = (uint64_t)(i == 123 ? (1 << 20) + 5 : 3000 + 17 * i);
            uint8_t* p = (uint8_t*)malloc(lens[i]);
            fill(p, lens[i], (uint32_t)(1000 + i));
            ptrs[i] = p;
        }
        const uint64_t launches0 = b200h_launch_count(ctx);
        rc = b200blob_upload_many(ctx, &t, ptrs, lens, N, ids, status, err);
        CHECK(rc == B200BLOB_E_MULTIPART);
        for (int i = 0; i < N; ++i) CHECK(status[i] == (i == 123 ? B200BLOB_E_MULTIPART : 0));
        CHECK(ids[123 * B200BLOB_ID_MAX] == 0 && strncmp(ids, "bl-", 3) == 0);
        printf("many_launches %" PRIu64 "\n", b200h_launch_count(ctx) - launches0);
        static char m[N * B200BLOB_MD5_B64_LEN], s[N * B200BLOB_SHA256_B64_LEN];
        CHECK(b200blob_hashes_many(ctx, ptrs, lens, N, m, s, err) == 0);
        for (int i = 0; i < N; i += 37)
            printf("hashes %u %" PRIu64 " %s %s -\n", (unsigned)(1000 + i), lens[i], m + i * B200BLOB_MD5_B64_LEN,
                   s + i * B200BLOB_SHA256_B64_LEN);
        for (int i = 0; i < N; ++i) free((void*)ptrs[i]);
    }
    b200h_destroy(ctx);
    printf("OK run\n");
    return 0;
}
