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


class DevicePayload:
    """A serialized function input that already lives in HBM: ``tensor`` is a contiguous CUDA tensor whose raw bytes ARE
    the wire payload (what a GPU-side serializer produced, or a tensor argument shipped raw).  A serializer hook may
    return one instead of ``bytes``; the map pump then digests it where it is (``batch.hash_table_tensors``: no
    host->device copy, nothing pickled through host memory) and copies it to the host once, for the PUT
    (the reference pickles tensor arguments on the CPU, py/modal/_utils/function_utils.py:577-621)."""

    __slots__ = ("tensor",)

    def __init__(self, tensor):
        if not (getattr(tensor, "is_cuda", False) and tensor.is_contiguous()):
            raise ValueError("DevicePayload needs a contiguous CUDA tensor")
        self.tensor = tensor

    def __len__(self) -> int:
        return self.tensor.numel() * self.tensor.element_size()

    def to_bytes(self) -> bytes:
        import torch

        flat = self.tensor.reshape(-1).view(torch.uint8)
        return flat.cpu().numpy().tobytes()


def hash_device_payloads(payloads: Sequence["DevicePayload"], ctx=None):
    """(sha256 uint8[n,32], md5 uint8[n,16]) numpy tables of payloads resident in HBM, hashed in place in one batch."""
    from . import batch

    d_sha, d_md5 = batch.hash_table_tensors([p.tensor for p in payloads], ctx=ctx)
    return d_sha.cpu().numpy(), d_md5.cpu().numpy()


def serialize_pickle(obj: Any) -> bytes:
    """cloudpickle protocol 4, like the reference's ``serialize`` (py/modal/_serialization.py:102-106)."""
    import cloudpickle

    return cloudpickle.dumps(obj, protocol=4)


def get_preferred_payload_format(payload_format: str | None = None) -> int:
    """``config.get("payload_format")`` of the reference (py/modal/_serialization.py:359-362; the setting is the
    ``MODAL_PAYLOAD_FORMAT`` environment variable or the config file): "cbor" selects CBOR, anything else pickle."""
    import os

    fmt = (payload_format or os.environ.get("MODAL_PAYLOAD_FORMAT") or "pickle").lower()
    return _wire.DATA_FORMAT_CBOR if fmt == "cbor" else _wire.DATA_FORMAT_PICKLE


def serialize_data_format(obj: Any, data_format: int) -> bytes:
    """The two argument formats of ``serialize_data_format`` (py/modal/_serialization.py:365-390)."""
    if data_format == _wire.DATA_FORMAT_PICKLE:
        return serialize_pickle(obj)
    if data_format == _wire.DATA_FORMAT_CBOR:
        try:
            import cbor2
        except ImportError:
            raise ExecutionError("CBOR support requires the 'cbor2' package to be installed.")
        try:
            return cbor2.dumps(obj)
        except cbor2.CBOREncodeError:
            typename = f"{type(obj).__module__}.{type(obj).__name__}"
            raise ExecutionError(f"Can not serialize type {typename} as cbor. If you need to use a custom data type, "
                                 "try to serialize it yourself e.g. by using pickle.dumps(my_data)")
    raise ExecutionError(f"Unknown data format {data_format!r}")


def _function_fields(function, payload_format: str | None = None) -> tuple[str | None, int, int]:
    """(method name, blob threshold, negotiated data format) as ``_create_input`` derives them (reference :589-601):
    the preferred format if the function's metadata lists it, else the first format it supports (pickle when the
    metadata lists none)."""
    meta = getattr(function, "_metadata", True)
    if not meta:
        raise ExecutionError("Attempted to call function that has not been hydrated with metadata")
    method_name = getattr(function, "_use_method_name", None) or None
    max_bytes = getattr(function, "_max_object_size_bytes", blob_utils.MAX_OBJECT_SIZE_BYTES)
    data_format = get_preferred_payload_format(payload_format)
    supported = list(getattr(meta, "supported_input_formats", None) or []) or [_wire.DATA_FORMAT_PICKLE]
    if data_format not in supported:
        data_format = supported[0]
    return method_name, max_bytes, data_format


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
                        serializer: Callable[[Any], bytes] | None = None, payload_format: str | None = None):
    """Serialize ``(args, kwargs)`` in the negotiated format and build the FunctionPutInputsItem, uploading to blob
    storage above the threshold (reference :576-621).  ``serializer`` overrides the format's serializer (tests)."""
    method_name, max_bytes, data_format = _function_fields(function, payload_format)
    payload = serializer((args, kwargs)) if serializer else serialize_data_format((args, kwargs), data_format)
    idx = idx or 0
    if should_upload(len(payload), max_bytes, function_call_invocation_type):
        upload = await blob_utils.blob_upload_with_r2_failure_info(payload, stub)
        return _blob_item(idx, upload, data_format, method_name)
    return _inline_item(idx, payload, data_format, method_name)


async def create_inputs_batch(argskwargs: Sequence[tuple[tuple, dict]], stub, *, function, first_idx: int = 0,
                              function_call_invocation_type=None,
                              serializer: Callable[[Any], bytes] | None = None, payload_format: str | None = None) -> list:
    """Items for inputs ``first_idx .. first_idx+len-1`` in order; all blobified payloads share one GPU batch."""
    method_name, max_bytes, data_format = _function_fields(function, payload_format)
    payloads = [serializer(ak) if serializer else serialize_data_format(ak, data_format) for ak in argskwargs]
    big = [i for i, p in enumerate(payloads) if should_upload(len(p), max_bytes, function_call_invocation_type)]
    uploads = dict(zip(big, await blob_utils.blob_upload_many([payloads[i] for i in big], stub))) if big else {}
    return [
        _blob_item(first_idx + i, uploads[i], data_format, method_name) if i in uploads
        else _inline_item(first_idx + i, p, data_format, method_name)
        for i, p in enumerate(payloads)
    ]
