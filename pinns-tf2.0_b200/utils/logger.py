"""Progress reporting for the B200 PINN core.

Keeps the call surface the reference scripts use (``Logger(hp)``; ``log_train_start/opt/epoch/end``,
``set_error_fn``, ``get_error_u``, ``get_elapsed``, ``get_epoch_duration`` -- utils/logger.py:7-60 of the reference) and
the layout of the lines it prints, but is written for an asynchronous device-resident trainer:

* the loss handed to ``log_train_epoch`` may be a lazy device scalar; it is converted to a float (one 8-byte read-back,
  one host sync) only for the epochs that are actually printed, i.e. every ``hp["log_frequency"]`` steps;
* under a multi-process launch (RANK/WORLD_SIZE in the environment) only rank 0 prints;
* there is no TensorFlow banner -- the native library reports its own version instead.
"""
import json
import os
import sys
import time


class _Stopwatch(object):
    """Wall-clock bookkeeping: total time since construction and time since the previous lap."""

    def __init__(self):
        self.t0 = self.lap_t = time.time()

    def total(self):
        return time.time() - self.t0

    def lap(self):
        now = time.time()
        dt, self.lap_t = now - self.lap_t, now
        return dt


def _micro(seconds):
    return int(round(seconds * 1e6))                        # datetime.fromtimestamp rounds to microseconds


def _mm_ss(seconds):
    s = _micro(seconds) // 1000000
    return "%02d:%02d" % ((s // 60) % 60, s % 60)          # the reference wraps at one hour too (quirk Q5)


def _ss_t(seconds):
    us = _micro(seconds)                                    # "%S.%f"[:-5] of the reference: tenths are TRUNCATED, not rounded
    return "%02d.%d" % ((us // 1000000) % 60, (us % 1000000) // 100000)


class Logger(object):
    def __init__(self, hp, stream=None):
        self.frequency = hp["log_frequency"]
        self.out = stream                                   # None: whatever sys.stdout is at print time, like print()
        self.quiet = int(os.environ.get("RANK", "0")) != 0
        self.watch = _Stopwatch()
        self.error_fn = None
        self.model = None
        self._say("Hyperparameters:")
        self._say(json.dumps(hp, indent=2))
        self._say("")
        self._say("PINN core: " + self._core_version())

    # ------------------------------------------------------------------ helpers
    def _say(self, text):
        if not self.quiet:
            print(text, file=self.out if self.out is not None else sys.stdout)

    @staticmethod
    def _core_version():
        try:
            import pinn_cabi
            return pinn_cabi.load().pinn_version().decode()
        except Exception as exc:      # the banner must never hide the real error, which surfaces at first use
            return "unavailable (%s)" % exc

    # compatibility attributes of the reference class
    @property
    def start_time(self):
        return self.watch.t0

    @property
    def prev_time(self):
        return self.watch.lap_t

    # ------------------------------------------------------------------ reference surface
    def get_elapsed(self):
        return _mm_ss(self.watch.total())

    def get_epoch_duration(self):
        return _ss_t(self.watch.lap())

    def set_error_fn(self, error_fn):
        self.error_fn = error_fn

    def get_error_u(self):
        return self.error_fn()

    def log_train_start(self, model, model_description=False):
        self.model = model
        self._say("\nTraining started")
        self._say("================")
        if model_description:
            self._say(model.summary())

    def log_train_opt(self, name):
        self._say("-- Starting %s optimization --" % name)

    def log_train_epoch(self, epoch, loss, custom="", is_iter=False):
        if epoch % self.frequency:
            return                                          # nothing printed: the (possibly lazy) loss is never fetched
        tag = "nt_epoch" if is_iter else "tf_epoch"
        self._say("%s = %6d  elapsed = %s (+%s)  loss = %.4e  %s"
                  % (tag, epoch, self.get_elapsed(), self.get_epoch_duration(), float(loss), custom))

    def log_train_end(self, epoch, custom=""):
        self._say("==================")
        self._say("Training finished (epoch %s): duration = %s  error = %.4e  %s"
                  % (epoch, self.get_elapsed(), self.get_error_u(), custom))
