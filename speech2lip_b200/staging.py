"""Per-frame input staging around the path.

The reference loads one `coords/<frame>.npy` ([500,500,2] float) per frame in the dataset (`src/data/someones_lip_dataset.py:
250-262`, `np.load` + `torch.Tensor`) and moves it to the GPU synchronously inside the render loop (`inference.py:141-142`,
`value.to(device)`), so at a few hundred frames/s the loop waits on the file system and on pageable H2D copies.
`NpyPrefetcher` keeps `depth` frames in flight: a background thread reads the next files into PINNED buffers and enqueues
their H2D copies on a copy stream; the consumer gets device tensors in file order and its stream waits on the copy's event
(no host synchronisation).  The audio windows and the uint8 BGR output staging are kernels (`renderer.audio_windows`,
`renderer.frames_to_bgr8`, `LipRenderer.render_sequence_host`); this class covers the file-backed inputs."""
import queue
import threading

import numpy as np
import torch


def _load_npy(path):
    try:
        return np.load(path)
    except ValueError:                      # the reference retries with allow_pickle=True (someones_lip_dataset.py:256-259)
        return np.load(path, allow_pickle=True)


class NpyPrefetcher:
    def __init__(self, paths, device, depth=4, dtype=torch.float32, loader=_load_npy):
        self.paths = list(paths)
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self.dtype = dtype
        self.loader = loader
        self._cuda = self.device.type == "cuda"
        self._q = queue.Queue(maxsize=self.depth)
        self._free = threading.Semaphore(self.depth)
        self._stop = False
        self._pinned = [None] * self.depth
        self._events = [None] * self.depth
        self._copy = torch.cuda.Stream(self.device) if self._cuda else None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def __len__(self):
        return len(self.paths)

    def _run(self):
        try:
            for i, path in enumerate(self.paths):
                self._free.acquire()
                if self._stop:
                    return
                slot = i % self.depth
                host = torch.from_numpy(np.ascontiguousarray(self.loader(path))).to(self.dtype)
                if not self._cuda:
                    self._q.put((host.clone(), None))
                    continue
                if self._events[slot] is not None:
                    self._events[slot].synchronize()          # the copy that last read this pinned buffer has finished
                buf = self._pinned[slot]
                if buf is None or buf.shape != host.shape:
                    buf = self._pinned[slot] = torch.empty(host.shape, dtype=self.dtype, pin_memory=True)
                buf.copy_(host)
                with torch.cuda.stream(self._copy):
                    dev = buf.to(self.device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy)
                self._events[slot] = ev
                self._q.put((dev, ev))
        except BaseException as e:                              # surfaced to the consumer instead of a silent short sequence
            self._q.put((e, None))

    def __iter__(self):
        for _ in range(len(self.paths)):
            item, ev = self._q.get()
            if isinstance(item, BaseException):
                raise RuntimeError("NpyPrefetcher: loading failed") from item
            if ev is not None:
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                item.record_stream(cur)                         # the caching allocator must not recycle it under the consumer
            self._free.release()
            yield item

    def close(self):
        self._stop = True
        self._free.release()
