"""CPU: scheduling logic of selfocc_b200.pipeline.FramePipeline with a fake torch.cuda (streams = ordered op logs).
Checks the three dependencies that make the overlap safe: compute k waits for upload k; an upload into a buffer waits
for the last frame that read it; download k waits for compute k; drain joins the download stream."""
from contextlib import contextmanager
import torch
from selfocc_b200.pipeline import FramePipeline


class _Event:
    def record(self, stream):
        self.stream = stream
        stream.ops.append(('record', self))


class _Stream:
    def __init__(self, name):
        self.name, self.ops = name, []

    def wait_event(self, ev):
        self.ops.append(('wait', ev))

    def wait_stream(self, s):
        self.ops.append(('wait_stream', s))


class _Cuda:
    def __init__(self):
        self.cur = self.main = _Stream('main')
        self.side = []

    def Stream(self, device=None):
        s = _Stream('side%d' % len(self.side))
        self.side.append(s)
        return s

    def Event(self):
        return _Event()

    def current_stream(self, device=None):
        return self.cur

    @contextmanager
    def stream(self, s):
        prev, self.cur = self.cur, s
        try:
            yield
        finally:
            self.cur = prev


def test_frame_pipeline_dependencies_and_data():
    cuda = _Cuda()
    frames = [[torch.full((4,), float(i)), torch.full((2, 3), 10.0 * i)] for i in range(5)]
    seen = []

    def compute(dev_in):
        cuda.cur.ops.append(('compute', len(seen)))
        seen.append((dev_in[0].clone(), dev_in[1].clone()))
        return {'a': dev_in[0] * 2, 'b': dev_in[1].sum().reshape(1)}

    out_host = [torch.zeros(4), torch.zeros(1)]
    pipe = FramePipeline(compute, lambda o: (o['a'], o['b']), out_host, torch.device('cpu'), cuda=cuda)
    h2d, d2h = cuda.side
    for k, f in enumerate(frames):
        pipe.submit(f, next_host=frames[k + 1] if k + 1 < len(frames) else None)
        assert torch.equal(out_host[0], f[0] * 2) and out_host[1].item() == f[1].sum().item()   # (synchronous on CPU)
    pipe.drain()
    # every frame was computed on its own data (double buffering never handed a stale / overwritten buffer to compute)
    for k, f in enumerate(frames):
        assert torch.equal(seen[k][0], f[0]) and torch.equal(seen[k][1], f[1])
    main = cuda.main.ops
    computes = [i for i, op in enumerate(main) if op[0] == 'compute']
    assert len(computes) == len(frames)
    done = []                                     # the event recorded right after each compute
    for i in computes:
        assert main[i + 1][0] == 'record'
        done.append(main[i + 1][1])
    uploads = [op[1] for op in h2d.ops if op[0] == 'record']     # `ready` events, one per upload, in frame order
    assert len(uploads) == len(frames)
    for k, i in enumerate(computes):             # compute k is preceded on the main stream by wait(ready_k)
        waits = [op[1] for op in main[(computes[k - 1] + 1 if k else 0):i] if op[0] == 'wait']
        assert uploads[k] in waits
    # upload k (k >= 2) reuses the buffer of frame k-2: it must wait for done[k-2] before copying
    h = h2d.ops
    rec_pos = [i for i, op in enumerate(h) if op[0] == 'record']
    for k in range(2, len(frames)):
        seg = h[rec_pos[k - 1] + 1:rec_pos[k]]
        assert ('wait', done[k - 2]) in seg
    # download k waits for done[k], in order; drain joins the download stream
    assert [op[1] for op in d2h.ops if op[0] == 'wait'] == done
    assert main[-1] == ('wait_stream', d2h)
    # uploads 1.. were issued BEFORE the compute of the previous frame was enqueued (that is the overlap)
    assert pipe.k == len(frames)
