"""The map input pump: raw (args, kwargs) -> FunctionPutInputsItem -> batched FunctionPutInputs.

Mirrors ``InputPreprocessor`` / ``InputPumper`` of the reference (py/modal/parallel_map.py:90-198) on the
part that touches the hash path.  The reference builds items one input at a time under an ordered 20-way
async map, which means one hashlib pass per payload on the event-loop thread; here the preprocessor drains
whatever is queued (up to ``HASH_WINDOW`` inputs) and hands the whole window to
``function_utils.create_inputs_batch`` -- one GPU hash batch per window -- preserving input order.
The retry / output-polling state machine of the reference (:1320-1644) is control plane and not rebuilt.
"""
from __future__ import annotations

import asyncio
from collections.abc import Callable
from typing import Any

from . import _wire
from ._logging import logger
from .function_utils import create_inputs_batch

MAP_INVOCATION_CHUNK_SIZE = 49  # inputs per FunctionPutInputs request (sync map)
SPAWN_MAP_INVOCATION_CHUNK_SIZE = 512
HASH_WINDOW = 1024  # inputs gathered into one GPU hash batch


class InputPreprocessor:
    def __init__(self, client, *, raw_input_queue, processed_input_queue: asyncio.Queue, function,
                 created_callback: Callable[[int], None] = lambda n: None,
                 done_callback: Callable[[], None] = lambda: None):
        self.client = client
        self.function = function
        self.inputs_created = 0
        self.raw_input_queue = raw_input_queue
        self.processed_input_queue = processed_input_queue
        self.created_callback = created_callback
        self.done_callback = done_callback

    async def _next_window(self) -> tuple[list, bool]:
        """Block for one raw input, then take everything else that is already queued (<= HASH_WINDOW)."""
        window, finished = [], False
        first = await self.raw_input_queue.get()
        if first is None:
            return window, True
        window.append(first)
        while len(window) < HASH_WINDOW:
            try:
                nxt = self.raw_input_queue.get_nowait()
            except asyncio.QueueEmpty:
                break
            if nxt is None:
                finished = True
                break
            window.append(nxt)
        return window, finished

    async def drain_input_generator(self):
        finished = False
        while not finished:
            window, finished = await self._next_window()
            if window:
                first_idx = self.inputs_created
                self.inputs_created += len(window)
                self.created_callback(self.inputs_created)
                items = await create_inputs_batch(window, self.client.stub, function=self.function, first_idx=first_idx)
                for item in items:
                    await self.processed_input_queue.put(item)
        await self.processed_input_queue.put(None)  # end-of-queue marker for the pumper
        self.done_callback()
        yield


async def queue_batch_iterator(q: asyncio.Queue, max_batch_size: int = 100, debounce_time: float = 0.015):
    """Lists of queued items (None ends the stream); flushes early when the queue runs dry
    (semantics of py/modal/_utils/async_utils.py:704-728)."""
    batch: list[Any] = []
    while True:
        if q.empty() and batch:
            yield batch
            batch = []
            await asyncio.sleep(debounce_time)
        item = await q.get()
        if len(batch) >= max_batch_size:
            yield batch
            batch = []
        if item is None:
            if batch:
                yield batch
            return
        batch.append(item)


class InputPumper:
    def __init__(self, client, *, input_queue: asyncio.Queue, function, function_call_id: str,
                 max_batch_size: int = MAP_INVOCATION_CHUNK_SIZE):
        self.client = client
        self.function = function
        self.input_queue = input_queue
        self.inputs_sent = 0
        self.function_call_id = function_call_id
        self.max_batch_size = max_batch_size

    async def pump_inputs(self):
        async for items in queue_batch_iterator(self.input_queue, max_batch_size=self.max_batch_size):
            request = _wire.FunctionPutInputsRequest(
                function_id=getattr(self.function, "object_id", ""), inputs=items, function_call_id=self.function_call_id)
            logger.debug(f"Pushing {len(items)} inputs to server. Queued: {self.input_queue.qsize()}.")
            await self.client.stub.FunctionPutInputs(request)
            self.inputs_sent += len(items)
        yield
