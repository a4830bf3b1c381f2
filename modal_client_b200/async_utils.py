"""The two asyncio helpers the upload path needs (cf. py/modal/_utils/async_utils.py:345-433,593-624)."""
from __future__ import annotations

import asyncio
import contextlib
import functools
import os


_ATTEMPT_TIMEOUT_FLOOR = 2.0  # an attempt is never given less than this, however little of the total budget is left


def retry(direct_fn=None, *, n_attempts: int = 3, base_delay: float = 0.0, delay_factor: float = 2.0,
          max_delay: float | None = None, attempt_timeout: float | None = 90, total_timeout: float | None = None):
    """Decorator: call an async function up to ``n_attempts`` times; the last failure propagates.

    Same knobs and behaviour as the reference's helper (py/modal/_utils/async_utils.py:345-433), which the upload path
    uses as ``@retry(n_attempts=3, base_delay=0.3, attempt_timeout=None)`` (blob_utils.py:110) and
    ``@retry(n_attempts=5, base_delay=0.1, attempt_timeout=None)`` (:377): between attempts it sleeps ``base_delay``,
    growing by ``delay_factor`` each time and capped at ``max_delay``; each attempt runs under
    ``asyncio.wait_for(attempt_timeout)`` (None: no limit); with ``total_timeout`` the attempt's limit also shrinks to
    what is left of the overall budget (never below 2 s) and a retry whose sleep would run past the budget is not
    made.  Cancellation is never retried.  ``RETRY_N_ATTEMPTS_OVERRIDE`` (environment; tests) replaces
    ``n_attempts``.  Usable bare (``@retry``) or with arguments."""
    import time

    def decorate(fn):
        @functools.wraps(fn)
        async def attempt_loop(*args, **kwargs):
            override = os.environ.get("RETRY_N_ATTEMPTS_OVERRIDE")
            attempts = int(override) if override else n_attempts
            deadline = None if total_timeout is None else time.time() + total_timeout
            pause = base_delay
            for attempt in range(attempts):
                limit = attempt_timeout
                if deadline is not None:
                    left = max(deadline - time.time(), _ATTEMPT_TIMEOUT_FLOOR)
                    limit = left if limit is None else min(limit, left)
                try:
                    if limit is None:
                        return await fn(*args, **kwargs)
                    return await asyncio.wait_for(fn(*args, **kwargs), timeout=limit)
                except asyncio.CancelledError:
                    raise
                except Exception:
                    if attempt == attempts - 1:
                        raise
                    if deadline is not None and time.time() + pause + _ATTEMPT_TIMEOUT_FLOOR >= deadline:
                        raise  # the budget would be gone after the sleep: fail now instead of sleeping first
                await asyncio.sleep(pause)
                pause *= delay_factor
                if max_delay is not None:
                    pause = min(pause, max_delay)

        return attempt_loop

    return decorate(direct_fn) if direct_fn is not None else decorate


@contextlib.asynccontextmanager
async def asyncnullcontext(*_args, **_kwargs):
    yield


async def gather_cancel_on_error(*coros):
    """Run coroutines concurrently; on the first failure cancel the rest and re-raise."""
    tasks = [asyncio.ensure_future(c) for c in coros]
    try:
        return await asyncio.gather(*tasks)
    except BaseException:
        for t in tasks:
            t.cancel()
        await asyncio.gather(*tasks, return_exceptions=True)
        raise


async def bounded_map(items, fn, concurrency: int) -> list:
    """``[await fn(x) for x in items]`` with at most ``concurrency`` calls in flight, results in input order.

    The reference drives its uploads with ``async_map(generator, fn, concurrency=N)`` (py/modal/_utils/async_utils.py),
    which keeps N coroutines alive however long the input is; this is the same bound for an input that is already a
    list -- a fixed set of worker tasks pulling indices -- instead of one task (plus a semaphore wait) per item,
    which for a million-file tree is a million pending tasks.  The first failure cancels the rest and re-raises."""
    items = list(items)
    n = len(items)
    results: list = [None] * n
    if n == 0:
        return results
    indices = iter(range(n))

    async def worker():
        for i in indices:  # one shared iterator: the event loop is single-threaded, every index is taken once
            results[i] = await fn(items[i])

    workers = [asyncio.ensure_future(worker()) for _ in range(max(1, min(concurrency, n)))]
    try:
        await asyncio.gather(*workers)
    except BaseException:
        for w in workers:
            w.cancel()
        await asyncio.gather(*workers, return_exceptions=True)
        raise
    return results


async def bounded_map_ordered(items, fn, concurrency: int):
    """Async generator: ``await fn(x)`` for every x of ``items`` (a list) with at most ``concurrency`` calls in flight,
    yielding the results IN INPUT ORDER as soon as each one and all its predecessors are finished -- the streaming
    behaviour of the reference's ``async_map_ordered`` (py/modal/_utils/async_utils.py:1200-1227) for an input that
    is already materialised.  At most ``2 * concurrency`` finished results wait for a slow predecessor (the
    reference's buffer bound).  The first failure cancels the rest and re-raises.  (Written for ~1 us of overhead per
    item: the map pump pushes 10^5 inputs through it.)"""
    items = list(items)
    n = len(items)
    if n == 0:
        return
    pending = object()
    results: list = [pending] * n
    next_in = 0
    next_out = 0
    error = None
    wake_consumer = asyncio.Event()
    wake_workers = asyncio.Event()
    window = max(1, 2 * concurrency)

    async def worker():
        nonlocal next_in, error
        while True:
            i = next_in
            if i >= n or error is not None:
                return
            if i - next_out >= window:  # too far ahead of the slowest predecessor: wait for the consumer
                wake_workers.clear()
                await wake_workers.wait()
                continue
            next_in = i + 1
            try:
                results[i] = await fn(items[i])
            except BaseException as exc:  # noqa: BLE001 - handed to the consumer
                if error is None:
                    error = exc
                wake_consumer.set()
                return
            if i == next_out:
                wake_consumer.set()

    workers = [asyncio.ensure_future(worker()) for _ in range(max(1, min(concurrency, n)))]
    try:
        while next_out < n:
            r = results[next_out]
            if r is not pending:  # finished results are handed out even if a later item has failed meanwhile
                results[next_out] = None
                next_out += 1
                wake_workers.set()
                yield r
                continue
            if error is not None:
                raise error
            wake_consumer.clear()
            await wake_consumer.wait()
    finally:
        for w in workers:
            w.cancel()
        await asyncio.gather(*workers, return_exceptions=True)


# items a worker may finish back to back before it lets the event loop run its other callbacks (with ``concurrency``
# workers taking turns, everything else on the loop waits for at most concurrency x this many items)
_YIELD_EVERY = 16


async def bounded_each_ordered(n: int, fn, concurrency: int, sink) -> None:
    """``sink(await fn(i))`` for i = 0..n-1 with at most ``concurrency`` calls in flight and ``sink`` called IN INDEX
    ORDER, each result as soon as it and all its predecessors are finished -- ``bounded_map_ordered`` with the consumer
    folded into the workers: whichever worker completes the oldest outstanding index hands the finished prefix on
    itself, so an item costs no task switch and no async-generator hop on its way out (the map pump pushes 10^5..10^6
    inputs through here on the event-loop thread).  ``sink`` is a plain callable; if it returns an awaitable (a bounded
    queue that is full) that is awaited before the next result is handed on.  At most ``2 * concurrency`` finished
    results wait behind a slow predecessor.  The first failure -- of ``fn`` or ``sink`` -- cancels the rest and
    re-raises; what the sink has seen by then is a prefix 0..k-1 (k <= the failing index) and nothing is handed on after it."""
    if n <= 0:
        return
    pending = object()
    results: list = [pending] * n
    next_in = 0
    next_out = 0
    draining = False
    error = None
    wake_workers = asyncio.Event()
    waiting_workers = 0
    window = max(1, 2 * concurrency)

    async def worker():
        nonlocal next_in, next_out, draining, error, waiting_workers
        streak = 0
        while True:
            i = next_in
            if i >= n or error is not None:
                return
            if i - next_out >= window:  # too far ahead of the slowest predecessor
                wake_workers.clear()
                waiting_workers += 1
                try:
                    await wake_workers.wait()
                finally:
                    waiting_workers -= 1
                continue
            next_in = i + 1
            try:
                r = await fn(i)
                if error is not None:  # somebody failed meanwhile: nothing is handed on after a failure
                    return
                if i != next_out or draining:
                    results[i] = r  # a predecessor is still out (or being handed on): whoever finishes it takes this along
                else:
                    draining = True  # one drainer at a time keeps the order while a sink call is being awaited
                    try:
                        while True:
                            next_out += 1
                            blocked = sink(r)
                            if blocked is not None:
                                await blocked
                                if error is not None:
                                    break
                            if next_out >= n:
                                break
                            r = results[next_out]
                            if r is pending:
                                break
                            results[next_out] = None
                    finally:
                        draining = False
                    if waiting_workers:
                        wake_workers.set()
            except BaseException as exc:  # noqa: BLE001 - re-raised by the caller below
                if error is None:
                    error = exc
                wake_workers.set()
                return
            streak += 1
            if streak >= _YIELD_EVERY:  # fn never suspended (cached / in-process stubs): do not starve the loop
                streak = 0
                await asyncio.sleep(0)

    workers = [asyncio.ensure_future(worker()) for _ in range(max(1, min(concurrency, n)))]
    try:
        await asyncio.gather(*workers)
    finally:
        for w in workers:
            w.cancel()
        await asyncio.gather(*workers, return_exceptions=True)
    if error is not None:
        raise error

