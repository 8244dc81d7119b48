"""Asynchronous frame output for render loops (reference render.py:520-548).

The reference's view loop finishes every frame on the critical path: `rgb.data.cpu()` (a device synchronisation), float -> uint8 on
the host, `imageio.imsave` -- about 25 ms per 480 x 270 frame.  That was noise next to a 20 s render; at 0.3 s per frame it is the
next bound, and on the 8-GPU path it is all that is left on rank 0.  `FrameSink` takes the frame off the render stream instead:

    quantise on the device (the reference's integerify: (img * 255).astype(uint8), render.py:508-509)  ->  copy to a pinned buffer on a
    side stream  ->  a worker thread waits for the copy's event, encodes the PNG and recycles the buffer

so the GPU starts view i+1 while view i is copied and encoded.  `render_views` is the reference's loop body on top of it.
The on-disk format is unchanged (8-bit RGB PNG, `{:05d}.png`).
"""
import os
import queue
import threading

import torch


class FrameSink:
    def __init__(self, out_dir, n_buffers=3, name_fmt='{:05d}.png', keep_frames=False):
        import cv2
        self._cv2 = cv2
        self.out_dir, self.name_fmt, self.keep = out_dir, name_fmt, keep_frames
        os.makedirs(out_dir, exist_ok=True)
        self._free = queue.Queue()
        self._work = queue.Queue()
        self._n = n_buffers
        self._bufs = {}                      # (H, W, C) -> pinned buffers are created on first use
        self._stream = None
        self.frames = {}
        self._err = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _buffer(self, shape):
        if shape not in self._bufs:
            self._bufs[shape] = True
            for _ in range(self._n):
                self._free.put(torch.empty(shape, dtype=torch.uint8).pin_memory())
        buf = self._free.get()               # blocks when all buffers are in flight: back-pressure on the render loop
        if tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, dtype=torch.uint8).pin_memory()
        return buf

    def submit(self, image, H, W, index):
        """image: CUDA float tensor [H*W, C] / [H, W, C] / [1, H*W, C] in [0, 1] (C = 3 or 1).  Returns at once."""
        if self._err is not None:
            raise self._err
        dev = image.device
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        q = (image.detach().reshape(H, W, -1) * 255.).to(torch.uint8)          # integerify on the device (truncation, like astype)
        buf = self._buffer(tuple(q.shape))
        ready = torch.cuda.Event()
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):
            buf.copy_(q, non_blocking=True)
            ready.record(self._stream)
        q.record_stream(self._stream)
        self._work.put((buf, ready, index))

    def _run(self):
        while True:
            item = self._work.get()
            if item is None:
                return
            buf, ready, index = item
            try:
                ready.synchronize()
                img = buf.numpy()
                if self.keep:
                    self.frames[index] = img.copy()
                out = img[..., ::-1] if img.shape[-1] == 3 else img[..., 0]      # OpenCV writes BGR
                if not self._cv2.imwrite(os.path.join(self.out_dir, self.name_fmt.format(index)), out):
                    raise IOError('could not write frame %s' % index)
            except Exception as e:           # surfaced by the next submit() / close()
                self._err = e
            finally:
                self._free.put(buf)

    def close(self):
        self._work.put(None)
        self._thread.join()
        if self._err is not None:
            raise self._err


def render_views(render_fn, c2ws, intrinsics, H, W, sink, start_index=1, **render_kwargs):
    """The body of the reference's view loop (render.py:520-548) with the frame output off the critical path: for every
    camera-to-world matrix, get_rays -> render_fn -> sink.submit(rgb).  `render_kwargs` are passed to render_fn unchanged
    (render.py passes show_progress, require_nablas, calc_normal, detailed_output=False and **render_kwargs_test).
    Returns the number of frames submitted; call sink.close() to wait for the files."""
    from . import rend_util
    dev = intrinsics.device
    n = 0
    for k, c2w in enumerate(c2ws):
        c2w_t = torch.as_tensor(c2w, dtype=torch.float32, device=dev)
        rays_o, rays_d, _ = rend_util.get_rays(c2w_t[None], intrinsics[None], H, W, N_rays=-1)
        with torch.no_grad():
            rgb, depth, extras = render_fn(rays_o, rays_d, **render_kwargs)
        sink.submit(rgb, H, W, start_index + k)
        n += 1
    return n
