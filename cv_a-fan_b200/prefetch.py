"""Input pipeline of the training loops: batch i+1 travels host -> device on a copy stream while the step of batch i runs.

The reference moves every batch with a blocking `.cuda()` at the top of the iteration (Classification/main_perturb.py:170-171,
Detection/train_aug_final.py:80-82); here the DataLoader's pinned batches are copied asynchronously one step ahead, so the
1.5 MB of a CIFAR batch (12.6 MB at 8 GPUs) is off the step's critical path.  Used by main_perturb.py and by bench.py's
end-to-end leg."""
import torch


class DevicePrefetcher:
    """Iterate a loader of tuples of (pinned) host tensors as tuples of device tensors, one batch ahead.
    None entries pass through.  Tensors handed out belong to the consumer's stream (record_stream)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        stream = torch.cuda.Stream(device=self.device)
        it = iter(self.loader)

        def fetch():
            try:
                batch = next(it)
            except StopIteration:
                return None
            with torch.cuda.stream(stream):
                dev = tuple(None if t is None else t.to(self.device, non_blocking=True) for t in batch)
                done = torch.cuda.Event()
                done.record(stream)
            return dev, done

        nxt = fetch()
        while nxt is not None:
            dev, done = nxt
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(done)
            for t in dev:
                if t is not None:
                    t.record_stream(cur)
            nxt = fetch()                    # enqueued BEFORE the consumer's step: overlaps it
            yield dev
