"""`training.trainer` of the drop-in: `OffSerialTrainer` with the reference's
constructor, `step()`, `train()` and `save_apprfunc()` (reference
training/trainer.py:15-158), reorganised around the GPU-resident engine:

* the six networks live on the GPU for the whole run (no `ModuleOnDevice`
  ping-pong, reference :64,:92); the CPU sampler and evaluator act with a CPU
  mirror of the policy that is refreshed from the flat parameter buffer every
  `policy_mirror_interval` iterations (1 = the reference's semantics: sample
  with the current weights);
* replay minibatches are gathered on the GPU (no per-step `.cuda()` copies, :72-74);
* `tb_info` is fetched from the device only on iterations that log.

Optional (kwargs, all off by default = the reference's serial semantics):
* `dsact_async_sampler=True` — the CPU env loop runs in a background thread and feeds the replay buffer
  asynchronously (north_star: "the off_sampler env loop stays on CPU and feeds the buffer asynchronously"); the
  trainer drains what has been collected at each iteration instead of waiting for `sample_batch_size` env steps,
  and publishes the policy without blocking (`publish_policy`: one async D2H copy; the sampler thread picks the
  snapshot up between two `sample()` calls);
* `dsact_full_checkpoint=True` — next to every `apprfunc_{it}.pkl` write `trainstate_{it}.pkl` (Adam moments,
  mean_std EMA, counters, generator state, replay ring); `dsact_resume_dir=<file>` restores it.
"""
__all__ = ["OffSerialTrainer", "create_trainer"]

import os
import queue
import threading
import time
from math import inf

import torch

from dsact_host import TB_TAGS as tb_tags


def _add_scalars(tb_info, writer, step):
    for key, value in tb_info.items():
        writer.add_scalar(key, value, step)


class OffSerialTrainer:
    def __init__(self, alg, sampler, buffer, evaluator, **kwargs):
        self.alg = alg
        self.sampler = sampler
        self.buffer = buffer
        self.evaluator = evaluator
        self.per_flag = kwargs["buffer_name"] == "prioritized_replay_buffer"
        if self.per_flag:
            raise NotImplementedError("prioritized replay is not part of the B200 update path")

        self.networks = self.alg.networks
        if kwargs.get("ini_network_dir") is not None:
            self.networks.load_state_dict(torch.load(kwargs["ini_network_dir"]))

        self.replay_batch_size = kwargs["replay_batch_size"]
        self.max_iteration = kwargs["max_iteration"]
        self.sample_interval = kwargs.get("sample_interval", 1)
        self.log_save_interval = kwargs["log_save_interval"]
        self.apprfunc_save_interval = kwargs["apprfunc_save_interval"]
        self.eval_interval = kwargs["eval_interval"]
        self.mirror_interval = max(1, int(kwargs.get("policy_mirror_interval", 1)))
        self.best_tar = -inf
        self.save_folder = kwargs["save_folder"]
        self.iteration = 0

        # the update engine is CUDA-only: the networks go to the GPU whatever `use_gpu` says
        self.use_gpu = True
        if not kwargs.get("use_gpu", False):
            print("dsac-v2_b200: the DSAC-T update runs on the CUDA engine; moving the networks to the GPU")
        self.networks.cuda()
        engine = self.networks.engine(self.replay_batch_size)
        if hasattr(self.buffer, "attach"):
            self.buffer.attach(engine)
            import weakref
            self.networks._attachments.append(weakref.ref(self.buffer))   # follows the engine if it is rebuilt for a larger batch

        # CPU mirror of the behaviour policy for sampler and evaluator (they were handed
        # `alg.networks` itself in the reference, :24-26)
        self.cpu_networks = self.sampler.networks
        self.evaluator.networks = self.cpu_networks
        self._policy_span = self._find_policy_span()
        self._policy_host = torch.empty(self._policy_span[1] - self._policy_span[0]).pin_memory()
        self.refresh_policy_mirror()

        self.writer = None
        if kwargs.get("dsact_tensorboard", True):
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=self.save_folder, flush_secs=20)
            _add_scalars({tb_tags["alg_time"]: 0, tb_tags["sampler_time"]: 0}, self.writer, 0)
            self.writer.flush()

        while self.buffer.size < kwargs["buffer_warm_size"]:
            samples, _ = self.sampler.sample()
            self.buffer.add_batch(samples)

        self.full_checkpoint = bool(kwargs.get("dsact_full_checkpoint", False))
        if kwargs.get("dsact_resume_dir"):
            self.load_trainstate(kwargs["dsact_resume_dir"])

        self.async_sampler = bool(kwargs.get("dsact_async_sampler", False))
        self._mirror_lock = threading.Lock()
        self._pub_lock, self._pub_event, self._pub_seq, self._applied_seq = threading.Lock(), None, 0, 0
        self._feed, self._stop, self._thread = queue.Queue(maxsize=64), threading.Event(), None
        if self.async_sampler:
            self._thread = threading.Thread(target=self._sampler_loop, name="dsact-sampler", daemon=True)
            self._thread.start()

        self.start_time = time.time()

    # ---- policy mirror ----------------------------------------------------------------
    def _find_policy_span(self):
        """[lo, hi) of the policy inside the engine's flat parameter buffer: after the critics (q1, q2 of DSAC-T; the
        single q of DSAC_V1), before log_alpha."""
        critics = [n for n in ("q1", "q2", "q") if hasattr(self.networks, n)]
        lo = sum(p.numel() for n in critics for p in getattr(self.networks, n).parameters())
        n_pi = sum(p.numel() for p in self.networks.policy.parameters())
        return lo, lo + n_pi

    def refresh_policy_mirror(self):
        """Copy the current policy weights GPU -> pinned host -> the CPU module sampler/evaluator use."""
        lo, hi = self._policy_span
        eng = self.networks.engine()
        lock = getattr(self, "_mirror_lock", None)   # absent during construction (no sampler thread yet)
        if lock:
            lock.acquire()                           # waits for a running sample() of the sampler thread
        try:
            with getattr(self, "_pub_lock", threading.Lock()):
                self._policy_host.copy_(eng.params[lo:hi], non_blocking=True)
                torch.cuda.current_stream(eng.device).synchronize()
                off = 0
                with torch.no_grad():
                    for p in self.cpu_networks.policy.parameters():
                        p.copy_(self._policy_host[off:off + p.numel()].view(p.shape))
                        off += p.numel()
                if hasattr(self, "_pub_seq"):
                    self._applied_seq = self._pub_seq    # anything published earlier is older than this snapshot
        finally:
            if lock:
                lock.release()

    def publish_policy(self):
        """Asynchronous mode: enqueue the GPU -> pinned host copy of the current policy and return; the sampler thread
        moves the snapshot into its CPU module between two `sample()` calls.  The trainer neither synchronises the
        stream nor waits for a running `sample()`."""
        lo, hi = self._policy_span
        eng = self.networks.engine()
        with self._pub_lock:   # held by the sampler thread only while it copies host -> module
            self._policy_host.copy_(eng.params[lo:hi], non_blocking=True)
            if self._pub_event is None:
                self._pub_event = torch.cuda.Event()
            self._pub_event.record(torch.cuda.current_stream(eng.device))
            self._pub_seq += 1

    def _apply_published(self):
        with self._pub_lock:
            if self._pub_seq == self._applied_seq:
                return
            self._pub_event.synchronize()
            off = 0
            with torch.no_grad():
                for p in self.cpu_networks.policy.parameters():
                    p.copy_(self._policy_host[off:off + p.numel()].view(p.shape))
                    off += p.numel()
            self._applied_seq = self._pub_seq

    # ---- asynchronous sampler feed --------------------------------------
    def _sampler_loop(self):
        while not self._stop.is_set():
            with self._mirror_lock:   # act with a consistent snapshot of the mirrored policy
                self._apply_published()
                samples, tb = self.sampler.sample()
            try:
                self._feed.put((samples, tb), timeout=1.0)
            except queue.Full:
                continue

    def _drain_feed(self):
        tb = {}
        while True:
            try:
                samples, tb = self._feed.get_nowait()
            except queue.Empty:
                return tb
            self.buffer.add_batch(samples)

    def close(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=5.0)
            self._thread = None

    # ---- full training state ------------------------------------------------------------------
    def save_trainstate(self, path):
        # "iteration" = the NEXT iteration to run: inside step() the state already holds this iteration's update
        nxt = self.iteration + (1 if getattr(self, "_in_step", False) else 0)
        torch.save({"alg": self.alg.full_state_dict(), "buffer": self.buffer.state_dict(), "iteration": nxt,
                    "best_tar": self.best_tar}, path)

    def load_trainstate(self, path):
        st = torch.load(path, weights_only=False)
        self.alg.load_full_state_dict(st["alg"])
        self.buffer.load_state_dict(st["buffer"])
        self.iteration, self.best_tar = int(st["iteration"]), st["best_tar"]
        self.refresh_policy_mirror()

    # ---- one iteration (reference :60-138) -----------------------------------------------
    def step(self):
        self._in_step = True
        try:
            self._step()
        finally:
            self._in_step = False

    def _step(self):
        sampler_tb_dict = {}
        if self.async_sampler:
            if self.iteration % self.mirror_interval == 0:
                self.publish_policy()
            sampler_tb_dict = self._drain_feed()
        elif self.iteration % self.sample_interval == 0:
            if self.iteration % self.mirror_interval == 0:
                self.refresh_policy_mirror()
            sampler_samples, sampler_tb_dict = self.sampler.sample()
            self.buffer.add_batch(sampler_samples)

        replay_samples = self.buffer.sample_batch(self.replay_batch_size)
        alg_tb_dict = self.alg.local_update(replay_samples, self.iteration)

        if self.iteration % self.log_save_interval == 0:
            print("Iter = ", self.iteration)
            if self.writer is not None:
                _add_scalars(alg_tb_dict, self.writer, step=self.iteration)
                _add_scalars(sampler_tb_dict, self.writer, step=self.iteration)
        self.last_tb = alg_tb_dict

        if self.iteration % self.eval_interval == 0:
            self.refresh_policy_mirror()
            total_avg_return = self.evaluator.run_evaluation(self.iteration)
            if total_avg_return >= self.best_tar and self.iteration >= self.max_iteration / 5:
                self.best_tar = total_avg_return
                print("Best return = {}!".format(str(self.best_tar)))
                folder = self.save_folder + "/apprfunc/"
                for filename in os.listdir(folder):
                    if filename.endswith("_opt.pkl"):
                        os.remove(folder + filename)
                torch.save(self.networks.state_dict(), folder + "apprfunc_{}_opt.pkl".format(self.iteration))
            if self.writer is not None:
                w, it = self.writer, self.iteration
                w.add_scalar(tb_tags["Buffer RAM of RL iteration"], self.buffer.__get_RAM__(), it)
                w.add_scalar(tb_tags["TAR of RL iteration"], total_avg_return, it)
                w.add_scalar(tb_tags["TAR of replay samples"], total_avg_return, it * self.replay_batch_size)
                w.add_scalar(tb_tags["TAR of total time"], total_avg_return, int(time.time() - self.start_time))
                w.add_scalar(tb_tags["TAR of collected samples"], total_avg_return,
                             self.sampler.get_total_sample_number())

        if self.iteration % self.apprfunc_save_interval == 0:
            self.save_apprfunc()

    def train(self):
        while self.iteration < self.max_iteration:
            self.step()
            self.iteration += 1
        self.save_apprfunc()
        self.close()
        if self.writer is not None:
            self.writer.flush()

    def save_apprfunc(self):
        os.makedirs(self.save_folder + "/apprfunc", exist_ok=True)
        torch.save(self.networks.state_dict(), self.save_folder + "/apprfunc/apprfunc_{}.pkl".format(self.iteration))
        if getattr(self, "full_checkpoint", False):
            self.save_trainstate(self.save_folder + "/apprfunc/trainstate_{}.pkl".format(self.iteration))


def create_trainer(alg, sampler, buffer, evaluator, **kwargs):
    trainer = OffSerialTrainer(alg, sampler, buffer, evaluator, **kwargs)
    print("Create trainer successfully!")
    return trainer
