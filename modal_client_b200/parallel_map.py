"""The map input pump: raw (args, kwargs) -> FunctionPutInputsItem -> batched FunctionPutInputs.

Mirrors ``InputPreprocessor`` / ``InputPumper`` of the reference (py/modal/parallel_map.py:90-198) and the
input-plane variant's ``create_input`` / ``drain_input_generator`` (:707-736) on the part that touches the hash path.
The reference builds items one input at a time under an ordered 20-way async map, which means one hashlib pass per
payload on the event-loop thread (``blob_upload_with_r2_failure_info`` -> ``get_upload_hashes``,
_utils/blob_utils.py:338-352).  Here the preprocessor is a three-stage pipeline that keeps the reference's
observable behaviour -- items leave in input order, each as soon as its own blob is uploaded, at most
``BLOB_MAX_PARALLELISM`` uploads in flight -- and feeds the GPU what it needs:

  collect   whatever is already queued, up to a BYTE budget (``HASH_WINDOW_BYTES``; one lone input is a window of
            one: nothing waits for company), serialized as it is taken;
  hash      all payloads of the window that must be blobified -> ONE GPU batch, on a worker thread
            (the library releases the GIL); up to ``HASH_WINDOWS_IN_FLIGHT`` windows on as many contexts, so
            packing + H2D of window k+1 overlaps the kernel of window k and the uploads of window k-1;
  upload    BlobCreate + PUT per payload, ``BLOB_MAX_PARALLELISM`` at a time, results emitted in order as they
            complete (``bounded_map_ordered``).

The retry / output-polling state machine of the reference (:1320-1644) is control plane and not rebuilt.
"""
from __future__ import annotations

import asyncio
import time
from collections.abc import Callable
from typing import Any

from . import _backend, _wire, blob_utils, hash_utils
from ._logging import logger
from .async_utils import bounded_each_ordered, bounded_map_ordered
from . import function_utils
from .function_utils import DevicePayload, _blob_item, _function_fields, _inline_item, serialize_data_format

MAP_INVOCATION_CHUNK_SIZE = 49  # inputs per FunctionPutInputs request (sync map)
SPAWN_MAP_INVOCATION_CHUNK_SIZE = 512
import os as _os

# payload bytes gathered into one GPU hash batch (B200H_PUMP_WINDOW_BYTES) ...
HASH_WINDOW_BYTES = int(_os.environ.get("B200H_PUMP_WINDOW_BYTES", 1 << 30))
HASH_WINDOW_ITEMS = 65536  # ... and at most this many inputs (both: of what is ALREADY queued)
# windows being hashed at the same time, one library context each (B200H_PUMP_WINDOWS_IN_FLIGHT)
HASH_WINDOWS_IN_FLIGHT = int(_os.environ.get("B200H_PUMP_WINDOWS_IN_FLIGHT", 2))
PUMP_INPUTS_MAX_RETRY_DELAY = 15.0  # reference :67 (RESOURCE_EXHAUSTED back-off ceiling of the pumper)

_END = object()


class _Window:
    """One hash window: the serialized payloads of consecutive inputs and which of them go to blob storage."""

    __slots__ = ("first_idx", "payloads", "big", "hashes", "big_payloads")

    def __init__(self, first_idx: int):
        self.first_idx = first_idx
        self.payloads: list[bytes] = []
        self.big: list[int] = []  # positions (within the window) of payloads above the threshold
        self.hashes = None  # future -> sequence of UploadHashes, one per entry of `big`
        self.big_payloads: list | None = None  # the payloads of `big` as handed to the hash call


class _WindowedInputPipeline:
    """collect -> hash -> upload over a raw-input queue; ``items()`` yields the wire items in input order."""

    def __init__(self, raw_input_queue, stub, function, *, first_idx: int = 0, on_created: Callable[[int], None] | None = None,
                 make_item: Callable[[Any], Any] | None = None, function_call_invocation_type=None,
                 serializer: Callable[[Any], bytes] | None = None, payload_format: str | None = None):
        self.q = raw_input_queue
        self.stub = stub
        self.method_name, self.max_bytes, self.data_format = _function_fields(function, payload_format)
        self.invocation_type = function_call_invocation_type
        self.serializer = serializer
        self.next_idx = first_idx
        self.first_idx = first_idx
        self.on_created = on_created or (lambda n: None)
        self.make_item = make_item  # None: the FunctionPutInputsItem itself
        self.windows_hashed = 0
        # where the time went (seconds): collecting+serializing, waiting for a free context, inside the GPU hash calls
        # (summed over worker threads), the upload stage waiting for a window's digests
        self.stats = {"collect_s": 0.0, "wait_context_s": 0.0, "hash_call_s": 0.0, "wait_hash_s": 0.0}
        self.digest_tables: list | None = None  # set to [] to keep every window's (sha[n,32], md5[n,16]) arrays

    # ---- stage 1: collect ----------------------------------------------------------------------------------
    async def _next_window(self) -> tuple[_Window | None, bool]:
        """Block for one raw input, then take what is already queued, up to the byte / item budget.  Every input is
        serialized as it is taken; ``should_upload()`` is written out (strictly greater than the function's limit; the
        8 KiB limit for async calls); ``on_created`` sees the running count after every input, like the reference's
        ``created_callback`` (:127)."""
        item = await self.q.get()
        if item is None:
            return None, True
        win = _Window(self.next_idx)
        payloads, big = win.payloads, win.big
        get_nowait, serializer, data_format, on_created = self.q.get_nowait, self.serializer, self.data_format, self.on_created
        max_bytes = self.max_bytes
        if self.invocation_type == _wire.FUNCTION_CALL_INVOCATION_TYPE_ASYNC:
            max_bytes = min(max_bytes, blob_utils.MAX_ASYNC_OBJECT_SIZE_BYTES)
        serialize = serializer if serializer else (lambda it: serialize_data_format(it, data_format))
        created = self.next_idx - self.first_idx
        nbytes, count, finished = 0, 0, False
        try:
            while True:
                payload = serialize(item)
                n = len(payload)
                if n > max_bytes:
                    big.append(count)
                payloads.append(payload)
                nbytes += n
                count += 1
                on_created(created + count)
                if nbytes >= HASH_WINDOW_BYTES or count >= HASH_WINDOW_ITEMS:
                    break
                try:
                    item = get_nowait()
                except asyncio.QueueEmpty:
                    break
                if item is None:
                    finished = True
                    break
        finally:
            self.next_idx += count
        return win, finished

    # ---- stage 2: hash -------------------------------------------------------------------------------------
    async def _collect_and_hash(self, out: asyncio.Queue):
        loop = asyncio.get_running_loop()
        pool = _backend.context_pool(HASH_WINDOWS_IN_FLIGHT)
        free = asyncio.Queue()
        for c in pool:
            free.put_nowait(c)
        finished = False
        try:
            def hash_window(c, payloads):
                t = time.perf_counter()
                try:
                    if set(map(type, payloads)) == {bytes}:  # (one C-level pass: this runs on a worker thread under the GIL)
                        return hash_utils.get_upload_hashes_many(payloads, ctx=c)
                    return self._hash_mixed_window(c, payloads)
                finally:
                    self.stats["hash_call_s"] += time.perf_counter() - t

            while not finished:
                t0 = time.perf_counter()
                win, finished = await self._next_window()
                self.stats["collect_s"] += time.perf_counter() - t0
                if win is None:
                    break
                if win.big:
                    t0 = time.perf_counter()
                    ctx = await free.get()  # at most one batch per context at a time
                    self.stats["wait_context_s"] += time.perf_counter() - t0
                    big_payloads = [win.payloads[i] for i in win.big]
                    win.big_payloads = big_payloads  # DevicePayloads are replaced by their host bytes in here
                    fut = loop.run_in_executor(None, hash_window, ctx, big_payloads)
                    fut.add_done_callback(lambda _f, c=ctx: free.put_nowait(c))
                    win.hashes = fut
                    self.windows_hashed += 1
                await out.put(win)
        finally:
            await out.put(_END)

    @staticmethod
    def _hash_mixed_window(ctx, payloads):
        """A window in which some payloads are ``DevicePayload``s: those are digested in HBM where they are (one
        batch), the host ones as usual (one batch); the device ones are copied to the host once, for their PUT.
        Returns a digest view in window order; ``payloads`` is updated in place with the materialised bytes."""
        import numpy as np

        dev = [i for i, p in enumerate(payloads) if type(p) is not bytes]
        host = [i for i, p in enumerate(payloads) if type(p) is bytes]
        n = len(payloads)
        sha, md5 = np.empty((n, 32), np.uint8), np.empty((n, 16), np.uint8)
        if host:
            s, m, _ = ctx.hash_buffers([payloads[i] for i in host], hash_utils.SHA256 | hash_utils.MD5)
            sha[host], md5[host] = s, m
        s, m = function_utils.hash_device_payloads([payloads[i] for i in dev], ctx=ctx)
        sha[dev], md5[dev] = s, m
        for i in dev:
            payloads[i] = payloads[i].to_bytes()  # the one device->host copy of this payload (its upload body)
        return hash_utils._UploadHashesView(sha, md5, n)

    # ---- stage 3: upload, emit in order ---------------------------------------------------------------------
    async def _hashed_windows(self):
        """(window, its UploadHashes sequence) in input order, each as soon as its hash batch is done."""
        handoff: asyncio.Queue = asyncio.Queue(maxsize=HASH_WINDOWS_IN_FLIGHT)
        producer = asyncio.ensure_future(self._collect_and_hash(handoff))
        try:
            while True:
                win = await handoff.get()
                if win is _END:
                    break
                t0 = time.perf_counter()
                hashes = await win.hashes if win.hashes is not None else ()
                self.stats["wait_hash_s"] += time.perf_counter() - t0
                if self.digest_tables is not None and win.hashes is not None:
                    self.digest_tables.append((hashes._sha, hashes._md5))
                yield win, hashes
            await producer  # surfaces a collector failure
        finally:
            if not producer.done():
                producer.cancel()
                await asyncio.gather(producer, return_exceptions=True)

    def _builder(self, win: _Window, hashes):
        """``build(pos)`` -> the wire item of the window's input ``pos`` (uploading its payload first if it is big)."""
        payloads, first_idx, stub = win.payloads, win.first_idx, self.stub
        data_format, method_name, make_item = self.data_format, self.method_name, self.make_item
        big_pos = {pos: k for k, pos in enumerate(win.big)}
        upload_row = blob_utils._blob_upload_row
        # the window's digest columns, sliced per row right here (no UploadHashes object per input)
        columns = getattr(hashes, "columns", None) if win.big else None
        if columns is not None and columns()[0] is None:
            columns = None  # a table without an MD5 column: rows go through the general function
        if win.big and columns is None:  # any other sequence of UploadHashes (a caller's own hash function)
            upload = blob_utils._blob_upload_bytes

            async def build_from_objects(pos):
                payload = payloads[pos]
                payloads[pos] = None
                k = big_pos.get(pos)
                if k is None:
                    item = _inline_item(first_idx + pos, payload, data_format, method_name)
                else:
                    item = _blob_item(first_idx + pos, await upload(hashes[k], payload, stub), data_format, method_name)
                return item if make_item is None else make_item(item)

            return build_from_objects
        md5_64, sha_64, md5_raw = columns() if columns else (None, None, None)

        async def build(pos):
            payload = payloads[pos]
            payloads[pos] = None  # the window must not pin every payload until its last upload is done
            k = big_pos.get(pos)
            if k is None:
                if type(payload) is not bytes:  # a small DevicePayload stays inline: one copy to the host
                    payload = payload.to_bytes()
                item = _inline_item(first_idx + pos, payload, data_format, method_name)
            else:
                if type(payload) is not bytes:
                    payload, win.big_payloads[k] = win.big_payloads[k], None  # materialised by the hash stage
                up = await upload_row(md5_64[24 * k : 24 * k + 24], sha_64[44 * k : 44 * k + 44],
                                      md5_raw[16 * k : 16 * k + 16], payload, stub)
                item = _blob_item(first_idx + pos, up, data_format, method_name)
            return item if make_item is None else make_item(item)

        return build

    async def run(self, emit: Callable[[Any], Any]) -> None:
        """Drive the pipeline to the end of the input; ``emit(item)`` receives every wire item in input order and may
        return an awaitable (a full bounded queue) that is then awaited."""
        async for win, hashes in self._hashed_windows():
            await bounded_each_ordered(len(win.payloads), self._builder(win, hashes), blob_utils.BLOB_MAX_PARALLELISM, emit)

    async def items(self):
        """The same as an async generator of wire items (the input-plane variant feeds a timestamped queue from it)."""
        async for win, hashes in self._hashed_windows():
            async for item in bounded_map_ordered(range(len(win.payloads)), self._builder(win, hashes),
                                                  blob_utils.BLOB_MAX_PARALLELISM):
                yield item


class InputPreprocessor:
    """Constructs FunctionPutInputsItem objects from the raw-input queue and puts them in the processed-input queue
    (reference :90-147)."""

    def __init__(self, client, *, raw_input_queue, processed_input_queue: asyncio.Queue, function,
                 created_callback: Callable[[int], None] = lambda n: None,
                 done_callback: Callable[[], None] = lambda: None,
                 serializer: Callable[[Any], bytes] | None = None):
        self.client = client
        self.function = function
        self.inputs_created = 0
        self.raw_input_queue = raw_input_queue
        self.processed_input_queue = processed_input_queue
        self.created_callback = created_callback
        self.done_callback = done_callback
        self.serializer = serializer  # None: the negotiated payload format's serializer (pickle / CBOR)
        self.hash_batches = 0
        self.stats: dict = {}
        self.keep_digest_tables = False  # True: ``digest_tables`` collects each window's (sha, md5) numpy tables
        self.digest_tables: list = []

    def _created(self, n: int) -> None:
        self.inputs_created = n
        self.created_callback(n)

    async def input_iter(self):
        """The raw inputs one by one until the end-of-input sentinel (reference :113-118)."""
        while True:
            raw_input = await self.raw_input_queue.get()
            if raw_input is None:
                break
            yield raw_input

    def create_input_factory(self):
        """The reference's per-input form (:120-135): ``await create_input((args, kwargs))`` numbers the input, reports
        it and builds its item on its own -- one hash launch per blobified input.  ``drain_input_generator`` does not
        use it (it batches); it is here for callers that drive inputs themselves."""

        async def create_input(argskwargs):
            idx = self.inputs_created
            self._created(idx + 1)
            args, kwargs = argskwargs
            return await function_utils._create_input(args, kwargs, self.client.stub, idx=idx, function=self.function,
                                                      serializer=(lambda ak: self.serializer(ak)) if self.serializer else None)

        return create_input

    async def drain_input_generator(self):
        pipe = _WindowedInputPipeline(self.raw_input_queue, self.client.stub, self.function, first_idx=0,
                                      on_created=self._created, serializer=self.serializer)
        if self.keep_digest_tables:
            pipe.digest_tables = self.digest_tables
        q = self.processed_input_queue
        # an unbounded queue (the reference's) never blocks: no fullness check, no coroutine round trip per item
        unbounded = getattr(q, "maxsize", 1) <= 0
        await pipe.run(q.put_nowait if unbounded else (lambda item: q.put_nowait(item) if not q.full() else q.put(item)))
        self.hash_batches = pipe.windows_hashed
        self.stats = pipe.stats
        await self.processed_input_queue.put(None)  # end-of-queue marker for the pumper
        self.done_callback()
        yield


async def queue_batch_iterator(q: asyncio.Queue, max_batch_size: int = 100, debounce_time: float = 0.015):
    """Lists of queued items (None ends the stream); flushes early when the queue runs dry
    (semantics of py/modal/_utils/async_utils.py:704-728: same batches, same debounce).  What is already queued is
    taken with ``get_nowait`` -- one coroutine round trip per batch instead of one per item."""
    batch: list[Any] = []
    get_nowait = getattr(q, "get_nowait", None)
    while True:
        if q.empty() and batch:
            yield batch
            batch = []
            await asyncio.sleep(debounce_time)
        item = await q.get()
        while True:
            if len(batch) >= max_batch_size:
                yield batch
                batch = []
            if item is None:
                if batch:
                    yield batch
                return
            batch.append(item)
            if get_nowait is None or q.empty():
                break
            item = get_nowait()


def _is_resource_exhausted(exc: BaseException) -> bool:
    """grpclib's ``GRPCError(Status.RESOURCE_EXHAUSTED)`` (status value 8), recognised structurally so that this
    module does not need grpclib installed."""
    status = getattr(exc, "status", None)
    return getattr(status, "name", None) == "RESOURCE_EXHAUSTED" or getattr(status, "value", status) == 8


class InputPumper:
    """Reads FunctionPutInputsItems from a queue and sends them to the server (reference :150-214), including the
    hooks of the map state machine: ``map_items_manager.add_items`` before the RPC (items are SENDING),
    ``handle_put_inputs_response`` after it (WAITING_FOR_OUTPUT with the server's input ids), and unlimited retries
    with capped exponential back-off while the server answers RESOURCE_EXHAUSTED (``_function_inputs_retry``)."""

    def __init__(self, client, *, input_queue: asyncio.Queue, function, function_call_id: str,
                 max_batch_size: int = MAP_INVOCATION_CHUNK_SIZE, map_items_manager=None):
        self.client = client
        self.function = function
        self.map_items_manager = map_items_manager
        self.input_queue = input_queue
        self.inputs_sent = 0
        self.function_call_id = function_call_id
        self.max_batch_size = max_batch_size
        self.resource_exhausted_retries = 0

    async def _put_inputs(self, request):
        delay = 0.1
        while True:
            try:
                return await self.client.stub.FunctionPutInputs(request)
            except Exception as exc:  # noqa: BLE001 - filtered below
                if not _is_resource_exhausted(exc):
                    raise
                self.resource_exhausted_retries += 1
                if self.resource_exhausted_retries % 8 == 0:
                    name = getattr(self.function, "_function_name", "")
                    logger.warning(f"Warning: map progress for function {name} is limited."
                                   " Common bottlenecks include slow iteration over results, or function backlogs.")
                await asyncio.sleep(delay)
                delay = min(delay * 2, PUMP_INPUTS_MAX_RETRY_DELAY)

    async def pump_inputs(self):
        assert self.client.stub
        async for items in queue_batch_iterator(self.input_queue, max_batch_size=self.max_batch_size):
            if self.map_items_manager is not None:
                await self.map_items_manager.add_items(items)
            request = _wire.FunctionPutInputsRequest(
                function_id=getattr(self.function, "object_id", ""), inputs=items, function_call_id=self.function_call_id)
            logger.debug(f"Pushing {len(items)} inputs to server. Num queued inputs awaiting push is {self.input_queue.qsize()}.")
            resp = await self._put_inputs(request)
            self.inputs_sent += len(items)
            if self.map_items_manager is not None:
                self.map_items_manager.handle_put_inputs_response(getattr(resp, "inputs", []))
        yield


# ------------------------------------------------------------------------------------------- input plane


class InputPlanePreprocessor:
    """The input half of ``_map_invocation_inputplane`` (reference :707-736): ``create_input`` numbers map calls from
    1, wraps each FunctionPutInputsItem into a ``MapStartOrContinueItem`` and ``drain_input_generator`` pushes them,
    in order, into the caller's timestamped queue; ``update_counters`` sees every created input and, at the end,
    ``set_have_all_inputs=True``.  Same windowed GPU pipeline underneath."""

    def __init__(self, client, *, raw_input_queue, queue, function,
                 update_counters: Callable[..., None] = lambda **kw: None,
                 serializer: Callable[[Any], bytes] | None = None):
        self.client = client
        self.function = function
        self.raw_input_queue = raw_input_queue
        self.queue = queue  # TimestampPriorityQueue-shaped: ``await queue.put(timestamp, item)``
        self.update_counters = update_counters
        self.serializer = serializer
        self.inputs_created = 0

    def _created(self, n: int) -> None:
        delta, self.inputs_created = n - self.inputs_created, n
        self.update_counters(created_delta=delta)

    async def drain_input_generator(self):
        pipe = _WindowedInputPipeline(
            self.raw_input_queue, self.client.stub, self.function, first_idx=1,  # 1-indexed map call idx (:708)
            on_created=self._created, serializer=self.serializer,
            make_item=lambda put_item: _wire.MapStartOrContinueItem(input=put_item))
        async for q_item in pipe.items():
            await self.queue.put(time.time(), q_item)
        self.update_counters(set_have_all_inputs=True)  # all inputs have been read
        yield
