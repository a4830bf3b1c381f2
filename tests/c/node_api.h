/*
 * node_api.h -- TEST STAND-IN for the subset of Node-API (https://nodejs.org/api/n-api.html, ABI-stable C API) that
 * bindings/node/b200hash_napi.c uses.  Node is not installed in this image; this header declares those functions
 * with Node's own signatures and tests/c/napi_host_check.c implements them over a tiny value model, so the addon is
 * compiled with -Wall -Werror and driven from C the way Node would drive it.  Not shipped, not a Node replacement.
 */
#ifndef B200H_TEST_NODE_API_H
#define B200H_TEST_NODE_API_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg = 1, napi_generic_failure = 9 } napi_status;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
#define NAPI_AUTO_LENGTH ((size_t)-1)
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_is_array(napi_env env, napi_value value, bool* result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_boolean(napi_env env, bool value, napi_value* result);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_value napi_register_module_v1(napi_env env, napi_value exports);
#endif
