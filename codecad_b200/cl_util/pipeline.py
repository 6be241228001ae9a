"""Host-side two-deep software pipelines of the reference, kept for callers that still
drive the per-block `.k.` kernels themselves (rendering/polygon2d.py, tests).

  interleave2  /root/reference/codecad/cl_util/__init__.py:8-67
  interleave   /root/reference/codecad/cl_util/cl_buffer.py:161-198

Same contract: at most two jobs in flight, LIFO work stack, a job is resumed after the
event it yielded has been waited for.  (subdivision() and mass_properties() in this
package do not use them any more: their whole hierarchy runs device-resident.)
"""


def interleave2(job_func, initial_jobs):
    stack = list(initial_jobs)
    pending = []  # [generator, event] of jobs that have yielded, oldest first; at most two

    def advance(gen):
        try:
            event = gen.send(None)
        except StopIteration as stop:
            if stop.value is not None:
                stack.extend(stop.value)
        else:
            pending.append([gen, event])

    while True:
        # top the pipeline up first: a new job runs to its first yield (its kernel is
        # enqueued) before we block on the older job's event
        while stack and len(pending) < 2:
            advance(job_func(stack.pop()))
        if not pending:
            return
        gen, event = pending.pop(0)
        event.wait()
        advance(gen)


def interleave(initial_jobs, helper1, helper2):
    assert len(initial_jobs), "There must be at least one job to start"
    stack = list(initial_jobs)
    busy, idle = helper1, helper2
    busy_event = busy.enqueue(*stack.pop())
    while True:
        idle_event = None
        if stack:
            idle_event = idle.enqueue(*stack.pop())
        stack.extend(busy.process_result(busy_event))
        if idle_event is None:
            if not stack:
                return
            idle_event = idle.enqueue(*stack.pop())
        busy, idle = idle, busy
        busy_event = idle_event
