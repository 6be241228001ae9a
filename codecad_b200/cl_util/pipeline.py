"""Host-side two-deep software pipelines of the reference, kept for callers that still
drive the per-block `.k.` kernels themselves (rendering/polygon2d.py, tests).

  interleave2  /root/reference/codecad/cl_util/__init__.py:8-67
  interleave   /root/reference/codecad/cl_util/cl_buffer.py:161-198

Same contract: at most two jobs in flight, LIFO work stack, a job is resumed after the
event it yielded has been waited for.  (subdivision() and mass_properties() in this
package do not use them any more: their whole hierarchy runs device-resident.)
"""


class _NoEvent:
    def wait(self):
        pass


def interleave2(job_func, initial_jobs):
    stack = list(initial_jobs)
    running = []  # [generator, pending event], oldest first, at most two
    while stack or running:
        while stack and len(running) < 2:
            running.append([job_func(stack.pop()), _NoEvent()])
        job = running.pop(0)
        job[1].wait()
        try:
            job[1] = job[0].send(None)
        except StopIteration as stop:
            if stop.value is not None:
                stack.extend(stop.value)
        else:
            running.append(job)


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
