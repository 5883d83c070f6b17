"""Result-directory convention of the reference (`utils/dirutils.py:73-128`), reduced to the entries the
hot path writes: `<dir_work>/result/<content>/<data_name>/<method>/<date>_<time>/<title>/{checkpoint,option,loss,...}`."""
import os


class Dir:
    SUBDIRS = ("checkpoint", "option", "loss", "log", "model", "train_img", "sample_img", "ema_sample_img")

    def __init__(self, task="train", content="test_code", dir_work="./", dir_dataset="", data_name="mnist", data_set="train",
                 data_size=64, date="", time="", method="base", title=""):
        root = os.path.join(dir_work, "result", content, data_name, method, f"{date}_{time}", title)
        self.list_dir = {}
        for s in self.SUBDIRS:
            p = os.path.join(root, s)
            os.makedirs(p, exist_ok=True)
            self.list_dir[s] = p
        self.root = root
