"""Request messages of the blob path.  With the real SDK installed these are the generated protobuf
classes (modal_proto/api.proto:811-817); standalone, structurally identical dataclasses stand in so the
upload functions can be driven by any stub object (tests use an in-process fake of BlobCreate)."""
from __future__ import annotations

import dataclasses

try:  # pragma: no cover - only with the real SDK present
    from modal_proto.api_pb2 import BlobCreateRequest, BlobGetRequest  # type: ignore
except Exception:

    @dataclasses.dataclass
    class BlobCreateRequest:  # type: ignore[no-redef]
        content_md5: str = ""
        content_sha256_base64: str = ""
        content_length: int = 0

    @dataclasses.dataclass
    class BlobGetRequest:  # type: ignore[no-redef]
        blob_id: str = ""


# enum values from modal_proto/api.proto:110-116,164-170
DATA_FORMAT_PICKLE = 1
DATA_FORMAT_CBOR = 4
FUNCTION_CALL_INVOCATION_TYPE_ASYNC = 3
FUNCTION_CALL_INVOCATION_TYPE_SYNC = 4

try:  # pragma: no cover
    from modal_proto.api_pb2 import FunctionInput, FunctionPutInputsItem, FunctionPutInputsRequest  # type: ignore
except Exception:

    @dataclasses.dataclass
    class FunctionInput:  # type: ignore[no-redef]  (api.proto:2156-2165)
        args: bytes | None = None
        args_blob_id: str | None = None
        data_format: int = DATA_FORMAT_PICKLE
        method_name: str | None = None
        final_input: bool = False

    @dataclasses.dataclass
    class FunctionPutInputsItem:  # type: ignore[no-redef]  (api.proto:2231-2237)
        idx: int = 0
        input: FunctionInput | None = None
        r2_failed: bool = False
        r2_throughput_bytes_s: int = 0

    @dataclasses.dataclass
    class FunctionPutInputsRequest:  # type: ignore[no-redef]
        function_id: str = ""
        inputs: list = dataclasses.field(default_factory=list)
        function_call_id: str = ""
