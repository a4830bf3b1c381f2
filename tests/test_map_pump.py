"""Map input pump + should_upload thresholds (host logic on the stand-in, and on the real GPU)."""
import asyncio
import hashlib
import pickle
import types

import pytest

from modal_client_b200 import _wire, blob_utils, function_utils, parallel_map
from tests.blob_server import FakeBlobStub, running_blob_server


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    return request.getfixturevalue("fake_backend" if request.param == "fake" else "gpu_backend")


def test_should_upload_thresholds():
    # py/test/should_upload_test.py:16-61 -- strict '>' on both limits; the async limit only for ASYNC calls
    M, A = blob_utils.MAX_OBJECT_SIZE_BYTES, blob_utils.MAX_ASYNC_OBJECT_SIZE_BYTES
    ASYNC, SYNC = _wire.FUNCTION_CALL_INVOCATION_TYPE_ASYNC, _wire.FUNCTION_CALL_INVOCATION_TYPE_SYNC
    assert not function_utils.should_upload(M, M, SYNC) and function_utils.should_upload(M + 1, M, SYNC)
    assert not function_utils.should_upload(A, M, ASYNC) and function_utils.should_upload(A + 1, M, ASYNC)
    assert not function_utils.should_upload(A + 1, M, SYNC) and not function_utils.should_upload(A + 1, M, None)
    assert function_utils.should_upload(M + 1, M, None)


def test_map_pump_blobifies_every_input_in_order(backend):
    # py/test/function_test.py:1555-1570 forces max_object_size_bytes=1 so every input becomes a blob
    fn = types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=1, _metadata=object(), object_id="fu-1")

    class Stub(FakeBlobStub):
        def __init__(self, host):
            super().__init__(host, multipart_threshold=50_000)
            self.put_batches = []

        async def FunctionPutInputs(self, req):
            self.put_batches.append(list(req.inputs))

    async def run():
        async with running_blob_server() as (host, store):
            stub = Stub(host)
            client = types.SimpleNamespace(stub=stub)
            raw, done = asyncio.Queue(), asyncio.Queue()
            inputs = [((i, "x" * (i * 977 % 70_000)), {"k": i}) for i in range(120)]
            for ak in inputs:
                raw.put_nowait(ak)
            raw.put_nowait(None)
            created = []
            pre = parallel_map.InputPreprocessor(client, raw_input_queue=raw, processed_input_queue=done, function=fn,
                                                 created_callback=created.append)
            pump = parallel_map.InputPumper(client, input_queue=done, function=fn, function_call_id="fc-1")

            async def drive(gen):
                async for _ in gen:
                    pass

            await asyncio.gather(drive(pre.drain_input_generator()), drive(pump.pump_inputs()))
            items = [it for b in stub.put_batches for it in b]
            assert [it.idx for it in items] == list(range(120)) and pump.inputs_sent == 120 and created[-1] == 120
            assert all(len(b) <= parallel_map.MAP_INVOCATION_CHUNK_SIZE for b in stub.put_batches)
            assert len(store.blobs) == 120  # one blob per input
            for it, ak in zip(items, inputs):
                assert it.input.args is None and pickle.loads(store.blobs[it.input.args_blob_id]) == ak
            # BlobCreate carried the GPU digests of exactly those payloads
            by_len = {r.content_length: r for r in stub.requests}
            p0 = function_utils.serialize_pickle(inputs[5])
            import base64

            assert by_len[len(p0)].content_sha256_base64 == base64.b64encode(hashlib.sha256(p0).digest()).decode()
            await blob_utils.ClientSessionRegistry.close_session()

    asyncio.run(run())


def test_small_inputs_stay_inline(backend):
    fn = types.SimpleNamespace(_use_method_name="m", _max_object_size_bytes=blob_utils.MAX_OBJECT_SIZE_BYTES, _metadata=1)

    async def run():
        items = await function_utils.create_inputs_batch([((1,), {}), ((2,), {})], stub=None, function=fn, first_idx=7)
        assert [i.idx for i in items] == [7, 8] and all(i.input.args and not i.input.args_blob_id for i in items)
        assert items[0].input.method_name == "m" and pickle.loads(items[1].input.args) == ((2,), {})
        one = await function_utils._create_input((3,), {}, None, function=fn, idx=2)
        assert one.idx == 2 and pickle.loads(one.input.args) == ((3,), {})

    asyncio.run(run())


def test_payload_format_negotiation(backend, monkeypatch):
    """_create_input's data-format choice (py/modal/_utils/function_utils.py:594-603): the preferred format when the
    function supports it, else the first supported one, pickle when the metadata lists none; CBOR payloads are cbor2."""
    import cbor2

    PICKLE, CBOR = _wire.DATA_FORMAT_PICKLE, _wire.DATA_FORMAT_CBOR

    def fn(supported):
        return types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=1 << 20,
                                     _metadata=types.SimpleNamespace(supported_input_formats=supported), object_id="fu-1")

    async def run():
        stub = types.SimpleNamespace()  # nothing is uploaded: the payloads stay inline
        item = await function_utils._create_input((1, "two"), {"k": [3]}, stub, function=fn([PICKLE, CBOR]), idx=7)
        assert item.idx == 7 and item.input.data_format == PICKLE and pickle.loads(item.input.args) == ((1, "two"), {"k": [3]})
        item = await function_utils._create_input((1, "two"), {"k": [3]}, stub, function=fn([PICKLE, CBOR]), payload_format="cbor")
        assert item.input.data_format == CBOR and cbor2.loads(item.input.args) == [[1, "two"], {"k": [3]}]
        item = await function_utils._create_input((1,), {}, stub, function=fn([PICKLE]), payload_format="cbor")
        assert item.input.data_format == PICKLE                      # preferred format unsupported -> first supported
        item = await function_utils._create_input((1,), {}, stub, function=fn([]), payload_format="cbor")
        assert item.input.data_format == PICKLE                      # nothing listed -> pickle
        monkeypatch.setenv("MODAL_PAYLOAD_FORMAT", "CBOR")
        items = await function_utils.create_inputs_batch([((i,), {}) for i in range(3)], stub, function=fn([CBOR, PICKLE]))
        assert [cbor2.loads(it.input.args) for it in items] == [[[i], {}] for i in range(3)]
        assert all(it.input.data_format == CBOR for it in items)
        with pytest.raises(function_utils.ExecutionError, match="as cbor"):
            await function_utils._create_input((object(),), {}, stub, function=fn([CBOR]))
        unhydrated = types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=1, _metadata=None)
        with pytest.raises(function_utils.ExecutionError, match="not been hydrated"):
            await function_utils._create_input((), {}, stub, function=unhydrated)

    asyncio.run(run())
