/*
 * napi_host_check.c -- drives bindings/node/b200hash_napi.c the way Node would: registers the module, builds an array
 * of Uint8Array payloads, calls exports.hashesMany(payloads) and exports.shouldUpload(n), prints what came back as
 * JSON.  The Node-API functions the addon calls are implemented here over a minimal value model (tests/c/node_api.h).
 * tests/test_c_host.py compares the digests with hashlib (js/src/blob.ts:35-36 semantics).
 */
#include <node_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum kind { K_ARRAY, K_U8, K_STRING, K_OBJECT, K_FUNCTION, K_NUMBER, K_BOOL };
struct napi_value__ {
    enum kind kind;
    uint32_t n;            /* array length / object property count */
    napi_value* items;     /* array elements / object values */
    char** names;          /* object keys */
    uint8_t* bytes;        /* Uint8Array data */
    size_t nbytes;
    char* str;
    napi_callback fn;
    double num;
    bool b;
};
struct napi_env__ { char error[512]; int threw; };
struct napi_callback_info__ { size_t argc; napi_value* argv; };

static napi_value mk(enum kind k) {
    napi_value v = (napi_value)calloc(1, sizeof *v);
    v->kind = k;
    return v;
}
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
    (void)env; (void)this_arg; (void)data;
    size_t want = *argc;
    for (size_t i = 0; i < want && i < cbinfo->argc; ++i) argv[i] = cbinfo->argv[i];
    *argc = cbinfo->argc;
    return napi_ok;
}
napi_status napi_is_array(napi_env env, napi_value value, bool* result) { (void)env; *result = value->kind == K_ARRAY; return napi_ok; }
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t* result) { (void)env; *result = value->n; return napi_ok; }
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value* result) {
    (void)env;
    if (object->kind != K_ARRAY || index >= object->n) return napi_invalid_arg;
    *result = object->items[index];
    return napi_ok;
}
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value) {
    (void)env;
    if (object->kind != K_ARRAY || index >= object->n) return napi_invalid_arg;
    object->items[index] = value;
    return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset) {
    (void)env; (void)arraybuffer; (void)byte_offset;
    if (v->kind != K_U8) return napi_invalid_arg;
    *type = napi_uint8_array;
    *length = v->nbytes;
    *data = v->bytes;
    return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value* result) { (void)env; *result = mk(K_OBJECT); return napi_ok; }
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value* result) {
    (void)env;
    napi_value a = mk(K_ARRAY);
    a->n = (uint32_t)length;
    a->items = (napi_value*)calloc(length ? length : 1, sizeof(napi_value));
    *result = a;
    return napi_ok;
}
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result) {
    (void)env;
    napi_value s = mk(K_STRING);
    size_t n = length == NAPI_AUTO_LENGTH ? strlen(str) : length;
    s->str = (char*)malloc(n + 1);
    memcpy(s->str, str, n);
    s->str[n] = 0;
    *result = s;
    return napi_ok;
}
napi_status napi_set_named_property(napi_env env, napi_value object, const char* name, napi_value value) {
    (void)env;
    object->items = (napi_value*)realloc(object->items, (object->n + 1) * sizeof(napi_value));
    object->names = (char**)realloc(object->names, (object->n + 1) * sizeof(char*));
    object->items[object->n] = value;
    object->names[object->n] = strdup(name);
    object->n += 1;
    return napi_ok;
}
napi_status napi_get_value_double(napi_env env, napi_value value, double* result) {
    (void)env;
    if (value->kind != K_NUMBER) return napi_invalid_arg;
    *result = value->num;
    return napi_ok;
}
napi_status napi_get_boolean(napi_env env, bool value, napi_value* result) { (void)env; *result = mk(K_BOOL); (*result)->b = value; return napi_ok; }
napi_status napi_create_function(napi_env env, const char* name, size_t length, napi_callback cb, void* data, napi_value* result) {
    (void)env; (void)name; (void)length; (void)data;
    *result = mk(K_FUNCTION);
    (*result)->fn = cb;
    return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char* code, const char* msg) {
    (void)code;
    env->threw = 1;
    snprintf(env->error, sizeof env->error, "%s", msg ? msg : "");
    return napi_ok;
}

static napi_value get(napi_value obj, const char* name) {
    for (uint32_t i = 0; i < obj->n; ++i)
        if (!strcmp(obj->names[i], name)) return obj->items[i];
    return NULL;
}
static napi_value call1(napi_env env, napi_value fn, napi_value arg) {
    struct napi_callback_info__ info = {1, &arg};
    return fn->fn(env, &info);
}

int main(int argc, char** argv) {
    /* payload sizes from the command line; contents: byte k of payload i = (i * 131 + k * 7 + 1) & 0xff */
    struct napi_env__ env_s = {{0}, 0};
    napi_env env = &env_s;
    napi_value exports = mk(K_OBJECT);
    napi_register_module_v1(env, exports);
    napi_value hashes_many = get(exports, "hashesMany"), should_upload = get(exports, "shouldUpload");
    if (!hashes_many || !should_upload) { fprintf(stderr, "exports incomplete\n"); return 2; }
    uint32_t n = (uint32_t)(argc - 1);
    napi_value arr;
    napi_create_array_with_length(env, n, &arr);
    for (uint32_t i = 0; i < n; ++i) {
        napi_value u8 = mk(K_U8);
        u8->nbytes = (size_t)strtoull(argv[i + 1], NULL, 10);
        u8->bytes = (uint8_t*)malloc(u8->nbytes ? u8->nbytes : 1);
        for (size_t k = 0; k < u8->nbytes; ++k) u8->bytes[k] = (uint8_t)(i * 131u + k * 7u + 1u);
        arr->items[i] = u8;
    }
    napi_value out = call1(env, hashes_many, arr);
    if (env->threw || !out) { printf("{\"error\": \"%s\"}\n", env->error); return 1; }
    napi_value md5 = get(out, "md5"), sha = get(out, "sha256");
    printf("{\"md5\": [");
    for (uint32_t i = 0; i < n; ++i) printf("%s\"%s\"", i ? ", " : "", md5->items[i]->str);
    printf("], \"sha256\": [");
    for (uint32_t i = 0; i < n; ++i) printf("%s\"%s\"", i ? ", " : "", sha->items[i]->str);
    napi_value num = mk(K_NUMBER);
    num->num = 2.0 * 1024 * 1024;
    bool at = call1(env, should_upload, num)->b;
    num->num += 1;
    bool above = call1(env, should_upload, num)->b;
    printf("], \"should_upload_2MiB\": %s, \"should_upload_2MiB_plus_1\": %s", at ? "true" : "false", above ? "true" : "false");
    /* wrong argument types must throw, not crash */
    env->threw = 0;
    call1(env, hashes_many, num);
    printf(", \"throws_on_non_array\": %s}\n", env->threw ? "true" : "false");
    return 0;
}
