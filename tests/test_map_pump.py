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


def test_every_blob_request_and_put_carries_the_digests_of_its_own_payload(backend, monkeypatch):
    """Inline and blobified inputs interleaved over several windows: the row of the window's digest table that goes
    into an input's BlobCreate request and its PUT must be that input's own (window position -> row among the big
    payloads -> row of the table), and the items must come out in input order with the right blob ids."""
    import base64

    fn = types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=3000, _metadata=object(), object_id="fu-2")
    monkeypatch.setattr(parallel_map, "HASH_WINDOW_BYTES", 40_000)  # ~20 windows

    class Stub:
        def __init__(self):
            self.requests = {}

        async def BlobCreate(self, req):
            blob_id = f"bl-{len(self.requests)}"
            self.requests[blob_id] = req
            if len(self.requests) % 3 == 0:
                await asyncio.sleep(0.001)  # uploads finish out of order
            return types.SimpleNamespace(WhichOneof=lambda _n: "upload_urls", blob_ids=[blob_id],
                                         upload_urls=types.SimpleNamespace(items=[f"null://{blob_id}"]))

        async def FunctionPutInputs(self, req):
            sent.extend(req.inputs)

    sent, puts = [], {}

    async def put(url, payload, content_md5_b64=None, content_type=None):
        puts[url.rsplit("/", 1)[1]] = (payload.data, content_md5_b64, payload.md5_checksum().hexdigest())
        return "etag"

    monkeypatch.setattr(blob_utils, "_upload_to_s3_url", put)
    rng = __import__("random").Random(5)
    # 2996 + 4 bytes = the limit itself: stays inline; 2997 + 4 is the first size that is blobified
    payloads = [bytes([i % 251]) * rng.choice([10, 2000, 2996, 2997, 5000, 9001]) + i.to_bytes(4, "little") for i in range(400)]

    async def run():
        stub = Stub()
        client = types.SimpleNamespace(stub=stub)
        raw, done = asyncio.Queue(), asyncio.Queue()
        for p in payloads:
            raw.put_nowait(p)
        raw.put_nowait(None)
        pre = parallel_map.InputPreprocessor(client, raw_input_queue=raw, processed_input_queue=done, function=fn,
                                             serializer=lambda p: p)
        pump = parallel_map.InputPumper(client, input_queue=done, function=fn, function_call_id="fc-2")

        async def drive(gen):
            async for _ in gen:
                pass

        await asyncio.gather(drive(pre.drain_input_generator()), drive(pump.pump_inputs()))
        assert pre.hash_batches >= 10
        assert [it.idx for it in sent] == list(range(len(payloads)))
        n_blobs = 0
        for it, p in zip(sent, payloads):
            if len(p) <= 3000:
                assert it.input.args == p and it.input.args_blob_id is None
                continue
            n_blobs += 1
            req = stub.requests[it.input.args_blob_id]
            body, md5_b64, md5_hex = puts[it.input.args_blob_id]
            assert body == p and req.content_length == len(p)
            assert req.content_md5 == md5_b64 == base64.b64encode(hashlib.md5(p).digest()).decode()
            assert req.content_sha256_base64 == base64.b64encode(hashlib.sha256(p).digest()).decode()
            assert md5_hex == hashlib.md5(p).hexdigest()
        assert n_blobs == len(stub.requests) == len(puts) and 100 < n_blobs < 300

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


# ------------------------------------------------------------------------------------- round 2: windowed pipeline


def test_bounded_map_ordered_streams_in_order_and_bounds_the_lookahead():
    from modal_client_b200.async_utils import bounded_map_ordered

    async def run():
        in_flight, peak, started = 0, 0, []

        async def fn(i):
            nonlocal in_flight, peak
            in_flight += 1
            peak = max(peak, in_flight)
            started.append(i)
            await asyncio.sleep(0.02 if i == 0 else 0.001)  # item 0 is slow: successors finish first
            in_flight -= 1
            return i * i

        out, seen_started_at_first_yield = [], None
        async for r in bounded_map_ordered(range(40), fn, concurrency=4):
            if seen_started_at_first_yield is None:
                seen_started_at_first_yield = len(started)
            out.append(r)
        assert out == [i * i for i in range(40)] and peak <= 4
        assert seen_started_at_first_yield <= 8 + 4  # at most 2 x concurrency results wait behind the slow one

        async def boom(i):
            if i == 5:
                raise RuntimeError("five")
            return i

        got = []
        with pytest.raises(RuntimeError, match="five"):
            async for r in bounded_map_ordered(range(20), boom, concurrency=3):
                got.append(r)
        assert got == list(range(5))
        assert [r async for r in bounded_map_ordered([], fn, 3)] == []

    asyncio.run(run())


def test_bounded_each_ordered_hands_results_on_in_order_without_a_consumer_task():
    """The map pump's upload stage: same contract as bounded_map_ordered (order, concurrency bound, look-ahead bound,
    first failure re-raised after the results before it), results delivered by calling ``sink`` -- which may answer
    with an awaitable (a full bounded queue) that must be awaited before the next result goes out."""
    from modal_client_b200 import async_utils
    from modal_client_b200.async_utils import bounded_each_ordered

    async def run():
        in_flight, peak, started = 0, 0, []

        async def fn(i):
            nonlocal in_flight, peak
            in_flight += 1
            peak = max(peak, in_flight)
            started.append(i)
            await asyncio.sleep(0.02 if i == 0 else 0.001)  # item 0 is slow: successors finish first
            in_flight -= 1
            return i * i

        out, started_at_first = [], []

        def sink(r):
            if not out:
                started_at_first.append(len(started))
            out.append(r)

        await bounded_each_ordered(40, fn, 4, sink)
        assert out == [i * i for i in range(40)] and peak <= 4
        assert started_at_first[0] <= 8 + 4  # at most 2 x concurrency results wait behind the slow one

        # a sink that blocks (bounded queue): order is kept although other workers finish while it is awaited
        q: asyncio.Queue = asyncio.Queue(maxsize=2)
        got = []

        async def consumer():
            while True:
                v = await q.get()
                if v is None:
                    return
                await asyncio.sleep(0.0005)
                got.append(v)

        async def quick(i):
            await asyncio.sleep(0)
            return i

        cons = asyncio.ensure_future(consumer())
        await bounded_each_ordered(60, quick, 5, lambda r: q.put_nowait(r) if not q.full() else q.put(r))
        await q.put(None)
        await cons
        assert got == list(range(60))

        async def boom(i):
            if i == 5:
                raise RuntimeError("five")
            return i

        seen = []
        with pytest.raises(RuntimeError, match="five"):
            await bounded_each_ordered(20, boom, 3, seen.append)
        assert seen == list(range(5))

        def bad_sink(r):
            if r == 3:
                raise ValueError("sink")
            seen2.append(r)

        seen2 = []
        with pytest.raises(ValueError, match="sink"):
            await bounded_each_ordered(10, quick, 2, bad_sink)
        assert seen2 == [0, 1, 2]
        await bounded_each_ordered(0, fn, 3, seen.append)  # nothing to do

        # fn that never suspends (in-process stubs) must not starve the loop: another task gets turns in between
        ticks = 0

        async def ticker():
            nonlocal ticks
            while True:
                await asyncio.sleep(0)
                ticks += 1

        async def instant(i):
            return i

        tk = asyncio.ensure_future(ticker())
        sunk = []
        await bounded_each_ordered(2000, instant, 20, sunk.append)
        tk.cancel()
        assert sunk == list(range(2000)) and ticks >= 2000 // (async_utils._YIELD_EVERY * 20) - 1 >= 4

    asyncio.run(run())


def test_queue_batch_iterator_batches_like_the_reference():
    """Same batches as py/modal/_utils/async_utils.py:704-728 for what is already queued: full batches while the queue
    holds enough, an early flush when it runs dry, None ends the stream."""

    async def reference_iter(q, max_batch_size=100, debounce_time=0.015):  # the reference's loop, restated
        item_list = []
        while True:
            if q.empty() and len(item_list) > 0:
                yield item_list
                item_list = []
                await asyncio.sleep(debounce_time)
            res = await q.get()
            if len(item_list) >= max_batch_size:
                yield item_list
                item_list = []
            if res is None:
                if len(item_list) > 0:
                    yield item_list
                break
            item_list.append(res)

    async def run():
        for n, bs in ((0, 4), (1, 4), (4, 4), (5, 4), (8, 4), (9, 4), (103, 49)):
            sizes = []
            for it in (reference_iter, parallel_map.queue_batch_iterator):
                q: asyncio.Queue = asyncio.Queue()
                for i in range(n):
                    q.put_nowait(i)
                q.put_nowait(None)
                batches = [b async for b in it(q, max_batch_size=bs, debounce_time=0)]
                assert [x for b in batches for x in b] == list(range(n))
                sizes.append([len(b) for b in batches])
            assert sizes[0] == sizes[1], (n, bs, sizes)

        # items that trickle in: a dry queue flushes what has been gathered
        q = asyncio.Queue()

        async def feed():
            for i in range(6):
                q.put_nowait(i)
                await asyncio.sleep(0.004)
            q.put_nowait(None)

        f = asyncio.ensure_future(feed())
        batches = [b async for b in parallel_map.queue_batch_iterator(q, max_batch_size=100, debounce_time=0.001)]
        await f
        assert [x for b in batches for x in b] == list(range(6)) and len(batches) >= 2

    asyncio.run(run())


def test_windows_are_byte_budgeted_and_items_stream_before_the_window_is_done(backend, monkeypatch):
    """the pump must not hold a whole window back -- an input reaches the processed queue as soon as its
    own upload (and its predecessors') is done -- and a window is capped by BYTES, not just by count."""
    fn = types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=1, _metadata=object(), object_id="fu-1")
    monkeypatch.setattr(parallel_map, "HASH_WINDOW_BYTES", 64 * 1024)
    uploads_done_when_first_item_arrived = []

    class Stub:
        def __init__(self):
            self.created = 0

        async def BlobCreate(self, req):
            self.created += 1
            return types.SimpleNamespace(WhichOneof=lambda _n: "upload_urls", blob_ids=["bl-x"],
                                         upload_urls=types.SimpleNamespace(items=["null://"]))

    put_calls = []

    async def slow_put(url, payload, content_md5_b64=None, content_type=None):
        put_calls.append(payload.size)
        await asyncio.sleep(0.002)
        return "etag"

    monkeypatch.setattr(blob_utils, "_upload_to_s3_url", slow_put)

    async def run():
        stub = Stub()
        client = types.SimpleNamespace(stub=stub)
        raw, done = asyncio.Queue(), asyncio.Queue()
        payloads = [bytes([i % 251]) * 4096 for i in range(100)]  # 400 KiB -> >= 6 windows of <= 64 KiB (+1 input)
        for p in payloads:
            raw.put_nowait(p)
        raw.put_nowait(None)
        created = []
        pre = parallel_map.InputPreprocessor(client, raw_input_queue=raw, processed_input_queue=done, function=fn,
                                             created_callback=created.append, serializer=lambda p: p)

        async def drive():
            async for _ in pre.drain_input_generator():
                pass

        task = asyncio.ensure_future(drive())
        first = await done.get()
        uploads_done_when_first_item_arrived.append(len(put_calls))
        items = [first]
        while (it := await done.get()) is not None:
            items.append(it)
        await task
        assert [it.idx for it in items] == list(range(100)) and created == list(range(1, 101))
        assert pre.hash_batches >= 6, pre.hash_batches
        assert uploads_done_when_first_item_arrived[0] < 40, "the first item waited for its whole window"
        assert stub.created == 100

    asyncio.run(run())


def test_pumper_map_items_manager_and_resource_exhausted_retry(monkeypatch):
    """InputPumper hooks of the reference (py/modal/parallel_map.py:174-214): add_items before the RPC,
    handle_put_inputs_response after it, unlimited retries while the server says RESOURCE_EXHAUSTED."""
    monkeypatch.setattr(parallel_map, "PUMP_INPUTS_MAX_RETRY_DELAY", 0.01)
    fn = types.SimpleNamespace(object_id="fu-1", _function_name="f")
    events = []

    class Exhausted(Exception):
        status = types.SimpleNamespace(name="RESOURCE_EXHAUSTED", value=8)

    class Manager:
        async def add_items(self, items):
            events.append(("add", [i.idx for i in items]))

        def handle_put_inputs_response(self, inputs):
            events.append(("resp", [i.idx for i in inputs]))

    class Stub:
        def __init__(self):
            self.calls = 0

        async def FunctionPutInputs(self, req):
            self.calls += 1
            if self.calls in (1, 2):
                raise Exhausted()
            if self.calls == 5:
                raise ValueError("not retried")
            return types.SimpleNamespace(inputs=[types.SimpleNamespace(idx=i.idx, input_id=f"in-{i.idx}") for i in req.inputs])

    async def run():
        q = asyncio.Queue()
        for i in range(5):
            q.put_nowait(_wire.FunctionPutInputsItem(idx=i))
        q.put_nowait(None)
        stub = Stub()
        pump = parallel_map.InputPumper(types.SimpleNamespace(stub=stub), input_queue=q, function=fn,
                                        function_call_id="fc", max_batch_size=2, map_items_manager=Manager())
        with pytest.raises(ValueError, match="not retried"):
            async for _ in pump.pump_inputs():
                pass
        assert pump.resource_exhausted_retries == 2 and pump.inputs_sent == 4
        assert events[:4] == [("add", [0, 1]), ("resp", [0, 1]), ("add", [2, 3]), ("resp", [2, 3])]

    asyncio.run(run())


def test_inputplane_preprocessor_wraps_items_and_counts(backend):
    """Input-plane variant (py/modal/parallel_map.py:707-736): 1-indexed map call idx, MapStartOrContinueItem wrapper,
    timestamped queue, created_delta per input and set_have_all_inputs at the end."""
    fn = types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=20_000, _metadata=object(), object_id="fu-1")

    class TsQueue:
        def __init__(self):
            self.items = []

        async def put(self, ts, item):
            assert isinstance(ts, float)
            self.items.append(item)

    async def run():
        async with running_blob_server() as (host, store):
            stub = FakeBlobStub(host)
            raw = asyncio.Queue()
            inputs = [((i, "y" * (3000 * i)), {}) for i in range(12)]  # the later ones cross the blob threshold
            for ak in inputs:
                raw.put_nowait(ak)
            raw.put_nowait(None)
            counters = []
            q = TsQueue()
            pre = parallel_map.InputPlanePreprocessor(types.SimpleNamespace(stub=stub), raw_input_queue=raw, queue=q,
                                                      function=fn, update_counters=lambda **kw: counters.append(kw))
            async for _ in pre.drain_input_generator():
                pass
            assert [it.input.idx for it in q.items] == list(range(1, 13))
            assert all(isinstance(it, _wire.MapStartOrContinueItem) and it.attempt_token is None for it in q.items)
            assert counters[:-1] == [{"created_delta": 1}] * 12 and counters[-1] == {"set_have_all_inputs": True}
            blobbed = [it for it in q.items if it.input.input.args_blob_id]
            assert 0 < len(blobbed) < 12
            for it, ak in zip(q.items, inputs):
                body = store.blobs[it.input.input.args_blob_id] if it.input.input.args_blob_id else it.input.input.args
                assert pickle.loads(body) == ak
            await blob_utils.ClientSessionRegistry.close_session()

    asyncio.run(run())


def test_bench_map_pump_leg_on_the_stand_in(fake_backend, monkeypatch):
    """bench.py's headline e2e leg (real InputPreprocessor/InputPumper, null control plane) on a tiny set: every
    BlobCreate carries the digests of exactly its payload, in order."""
    import base64

    import bench

    monkeypatch.setattr(blob_utils, "_upload_to_s3_url", bench._null_put)
    payloads = [bytes([i]) * (1000 + 37 * i) for i in range(50)]
    stub = bench.NullStub()
    dt, batches, tables = bench.run_map_pump(payloads, stub)
    assert dt > 0 and batches >= 1 and stub.inputs_put == 50 and len(stub.blob_requests) == 50
    for p, r in zip(payloads, stub.blob_requests):
        assert r.content_length == len(p)
        assert r.content_sha256_base64 == base64.b64encode(hashlib.sha256(p).digest()).decode()
        assert r.content_md5 == base64.b64encode(hashlib.md5(p).digest()).decode()
    assert sum(len(t[0]) for t in tables) == 50


def _run_pump_with_device_payloads(make_device_payload, raw_bytes):
    """Inputs alternate between ordinary bytes and payloads that live 'in HBM'; every one is above the blob threshold
    except the last.  Returns (items, blob store, BlobCreate requests)."""
    fn = types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=5_000, _metadata=object(), object_id="fu-1")
    out = {}

    async def run():
        async with running_blob_server() as (host, store):
            stub = FakeBlobStub(host)
            raw, done = asyncio.Queue(), asyncio.Queue()
            for i, b in enumerate(raw_bytes):
                raw.put_nowait(make_device_payload(b) if i % 2 == 0 else b)
            raw.put_nowait(None)
            pre = parallel_map.InputPreprocessor(types.SimpleNamespace(stub=stub), raw_input_queue=raw,
                                                 processed_input_queue=done, function=fn, serializer=lambda p: p)
            async for _ in pre.drain_input_generator():
                pass
            items = []
            while (it := await done.get()) is not None:
                items.append(it)
            out.update(items=items, blobs=dict(store.blobs), requests=list(stub.requests))
            await blob_utils.ClientSessionRegistry.close_session()

    asyncio.run(run())
    return out["items"], out["blobs"], out["requests"]


def _check_device_payload_run(items, blobs, requests, raw_bytes):
    import base64

    assert [it.idx for it in items] == list(range(len(raw_bytes)))
    by_len = {r.content_length: r for r in requests}
    for it, b in zip(items, raw_bytes):
        if len(b) > 5_000:
            assert blobs[it.input.args_blob_id] == b  # the PUT body is the payload, wherever it lived
            r = by_len[len(b)]
            assert r.content_sha256_base64 == base64.b64encode(hashlib.sha256(b).digest()).decode()
            assert r.content_md5 == base64.b64encode(hashlib.md5(b).digest()).decode()
        else:
            assert it.input.args == b and not it.input.args_blob_id


def test_pump_takes_payloads_that_live_on_the_device(fake_backend, monkeypatch):
    """SURVEY 8(f)4 wiring, host logic: a serializer may hand the pump a DevicePayload; it is digested by
    function_utils.hash_device_payloads (in HBM on the real thing), copied to the host once for its PUT, and mixes
    freely with ordinary bytes payloads in one window."""
    from oracle import c_oracle

    class HostStandIn:  # same surface as function_utils.DevicePayload, bytes kept on the host for the CPU test
        def __init__(self, b):
            self.b, self.copies = b, 0

        def __len__(self):
            return len(self.b)

        def to_bytes(self):
            self.copies += 1
            return self.b

    made = []

    def make(b):
        made.append(HostStandIn(b))
        return made[-1]

    def fake_hash_device_payloads(payloads, ctx=None):
        import numpy as np

        return (np.frombuffer(b"".join(c_oracle.sha256(p.b) for p in payloads), np.uint8).reshape(-1, 32),
                np.frombuffer(b"".join(c_oracle.md5(p.b) for p in payloads), np.uint8).reshape(-1, 16))

    monkeypatch.setattr(function_utils, "hash_device_payloads", fake_hash_device_payloads)
    raw_bytes = [bytes([i + 1]) * (6_000 + 977 * i) for i in range(9)] + [b"tiny"]
    items, blobs, requests = _run_pump_with_device_payloads(make, raw_bytes)
    _check_device_payload_run(items, blobs, requests, raw_bytes)
    assert all(p.copies == 1 for p in made)  # one device->host copy each, never more


@pytest.mark.gpu
def test_pump_hashes_cuda_tensor_payloads_in_hbm(gpu_backend):
    """The same on the GPU: CUDA tensors as payloads, digested in place by batch.hash_table_tensors."""
    import torch

    raw_bytes = [bytes([i + 1]) * (6_000 + 977 * i) for i in range(9)] + [b"tiny"]
    launches0 = gpu_backend.launch_count

    def make(b):
        return function_utils.DevicePayload(torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda())

    items, blobs, requests = _run_pump_with_device_payloads(make, raw_bytes)
    _check_device_payload_run(items, blobs, requests, raw_bytes)
    assert gpu_backend.launch_count > launches0


def test_per_input_form_of_the_preprocessor(backend):
    """``input_iter`` / ``create_input_factory`` of the reference's InputPreprocessor (py/modal/parallel_map.py:113-135):
    inputs are numbered from 0 in call order, ``created_callback`` sees the running count, small inputs stay inline and
    big ones are blobified -- the same items the batched ``drain_input_generator`` produces."""
    fn = types.SimpleNamespace(_use_method_name="", _max_object_size_bytes=2000, _metadata=object(), object_id="fu-3")

    async def run():
        async with running_blob_server() as (host, store):
            client = types.SimpleNamespace(stub=FakeBlobStub(host, multipart_threshold=10**9))
            raw = asyncio.Queue()
            inputs = [((i,), {"pad": "y" * (i * 900)}) for i in range(5)]
            for ak in inputs:
                raw.put_nowait(ak)
            raw.put_nowait(None)
            created = []
            pre = parallel_map.InputPreprocessor(client, raw_input_queue=raw, processed_input_queue=asyncio.Queue(),
                                                 function=fn, created_callback=created.append)
            create_input = pre.create_input_factory()
            items = [await create_input(ak) async for ak in pre.input_iter()]
            assert [it.idx for it in items] == [0, 1, 2, 3, 4] and created == [1, 2, 3, 4, 5] and pre.inputs_created == 5
            for it, ak in zip(items, inputs):
                blob = it.input.args_blob_id
                assert pickle.loads(store.blobs[blob] if blob else it.input.args) == ak
            assert [bool(it.input.args_blob_id) for it in items] == [False, False, False, True, True]
            await blob_utils.ClientSessionRegistry.close_session()

    asyncio.run(run())
