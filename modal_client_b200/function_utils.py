"""Input construction for function calls: serialize -> size test -> (blobify) -> FunctionPutInputsItem.

Counterpart of py/modal/_utils/function_utils.py:562-625.  ``_create_input`` keeps the reference's
one-at-a-time shape; ``create_inputs_batch`` is what the map pump calls here: it serializes a window of
inputs, sends every payload that must be blobified to the GPU in ONE hash batch
(``blob_utils.blob_upload_many``), and returns the items in input order.
"""
from __future__ import annotations

from collections.abc import Callable, Sequence
from typing import Any

from . import _wire, blob_utils
from .exception import ExecutionError


def should_upload(num_bytes: int, max_object_size_bytes: int, function_call_invocation_type=None) -> bool:
    """Strictly-greater-than on both limits (reference :562-573; pinned by py/test/should_upload_test.py)."""
    if num_bytes > max_object_size_bytes:
        return True
    return (
        function_call_invocation_type == _wire.FUNCTION_CALL_INVOCATION_TYPE_ASYNC
        and num_bytes > blob_utils.MAX_ASYNC_OBJECT_SIZE_BYTES
    )


def serialize_pickle(obj: Any) -> bytes:
    """cloudpickle protocol 4, like the reference's ``serialize`` (py/modal/_serialization.py:102-106)."""
    import cloudpickle

    return cloudpickle.dumps(obj, protocol=4)


def _function_fields(function) -> tuple[str | None, int, int]:
    if getattr(function, "_metadata", True) is None:
        raise ExecutionError("Attempted to call function that has not been hydrated with metadata")
    method_name = getattr(function, "_use_method_name", None) or None
    max_bytes = getattr(function, "_max_object_size_bytes", blob_utils.MAX_OBJECT_SIZE_BYTES)
    return method_name, max_bytes, _wire.DATA_FORMAT_PICKLE


def _inline_item(idx: int, payload: bytes, data_format: int, method_name) -> "_wire.FunctionPutInputsItem":
    return _wire.FunctionPutInputsItem(
        idx=idx, input=_wire.FunctionInput(args=payload, data_format=data_format, method_name=method_name))


def _blob_item(idx: int, upload: tuple[str, bool, int], data_format: int, method_name):
    blob_id, r2_failed, r2_bps = upload
    return _wire.FunctionPutInputsItem(
        idx=idx,
        input=_wire.FunctionInput(args_blob_id=blob_id, data_format=data_format, method_name=method_name),
        r2_failed=r2_failed,
        r2_throughput_bytes_s=r2_bps,
    )


async def _create_input(args, kwargs, stub, *, function, idx: int | None = None, function_call_invocation_type=None,
                        serializer: Callable[[Any], bytes] = serialize_pickle):
    method_name, max_bytes, data_format = _function_fields(function)
    payload = serializer((args, kwargs))
    idx = idx or 0
    if should_upload(len(payload), max_bytes, function_call_invocation_type):
        upload = await blob_utils.blob_upload_with_r2_failure_info(payload, stub)
        return _blob_item(idx, upload, data_format, method_name)
    return _inline_item(idx, payload, data_format, method_name)


async def create_inputs_batch(argskwargs: Sequence[tuple[tuple, dict]], stub, *, function, first_idx: int = 0,
                              function_call_invocation_type=None,
                              serializer: Callable[[Any], bytes] = serialize_pickle) -> list:
    """Items for inputs ``first_idx .. first_idx+len-1`` in order; all blobified payloads share one GPU batch."""
    method_name, max_bytes, data_format = _function_fields(function)
    payloads = [serializer(ak) for ak in argskwargs]
    big = [i for i, p in enumerate(payloads) if should_upload(len(p), max_bytes, function_call_invocation_type)]
    uploads = dict(zip(big, await blob_utils.blob_upload_many([payloads[i] for i in big], stub))) if big else {}
    return [
        _blob_item(first_idx + i, uploads[i], data_format, method_name) if i in uploads
        else _inline_item(first_idx + i, p, data_format, method_name)
        for i, p in enumerate(payloads)
    ]
