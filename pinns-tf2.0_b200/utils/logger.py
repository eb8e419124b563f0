"""Progress logger with the reference's surface (utils/logger.py:7-60), TensorFlow-free.

Same constructor (``Logger(hp)`` reading hp["log_frequency"]) and methods; the TF banner (logger.py:13-15) is
replaced by the native library's version line.  The loss handed to ``log_train_epoch`` may be a lazy device
scalar: it is only fetched (one 8-byte D2H) when a line is actually printed, i.e. every ``frequency`` epochs.
"""
import json
import time
from datetime import datetime


class Logger(object):
    def __init__(self, hp):
        print("Hyperparameters:")
        print(json.dumps(hp, indent=2))
        print()
        try:
            import pinn_cabi
            print("PINN core: {}".format(pinn_cabi.load().pinn_version().decode()))
        except Exception as e:  # the banner must not hide the real error, which surfaces at first use
            print("PINN core: unavailable ({})".format(e))
        self.start_time = time.time()
        self.prev_time = self.start_time
        self.frequency = hp["log_frequency"]

    def get_epoch_duration(self):
        now = time.time()
        edur = datetime.fromtimestamp(now - self.prev_time).strftime("%S.%f")[:-5]
        self.prev_time = now
        return edur

    def get_elapsed(self):
        return datetime.fromtimestamp(time.time() - self.start_time).strftime("%M:%S")

    def get_error_u(self):
        return self.error_fn()

    def set_error_fn(self, error_fn):
        self.error_fn = error_fn

    def log_train_start(self, model, model_description=False):
        print("\nTraining started")
        print("================")
        self.model = model
        if model_description:
            print(model.summary())

    def log_train_epoch(self, epoch, loss, custom="", is_iter=False):
        if epoch % self.frequency == 0:
            name = "nt_epoch" if is_iter else "tf_epoch"
            print(f"{name} = {epoch:6d}  elapsed = {self.get_elapsed()} (+{self.get_epoch_duration()})  "
                  f"loss = {float(loss):.4e}  " + custom)

    def log_train_opt(self, name):
        print(f"-- Starting {name} optimization --")

    def log_train_end(self, epoch, custom=""):
        print("==================")
        print(f"Training finished (epoch {epoch}): duration = {self.get_elapsed()}  "
              f"error = {self.get_error_u():.4e}  " + custom)
