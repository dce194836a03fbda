"""A small stand-in with the reference Network's call surface: the reference
itself does not travel to the GPU box, so the drop-in ``forward`` functions are exercised on this.

Same attributes/methods the replaced ``forward``s touch: ``maxdisp``, ``disp``,
``feature(img, task_arch, path)``, ``matching(cost, task_arch, path)``, ``search_feature(img, ops)``,
``search_matching(cost, ops, t)``; per-layer ``nn.ModuleList``s indexed by ``task_arch[name][0]`` like
rag_model.py:291-295,331-336."""
import torch
import torch.nn as nn


class MirrorNet(nn.Module):
    def __init__(self, disp_module, maxdisp=192, n_paths=2, c=12):
        super().__init__()
        self.maxdisp = maxdisp
        self.disp = disp_module
        self.stem2d = nn.ModuleList([nn.Conv2d(3, c, 3, stride=3, padding=0, bias=False) for _ in range(n_paths)])
        self.last2d = nn.ModuleList([nn.Conv2d(c, c, 1, bias=False) for _ in range(n_paths)])
        self.stem3d = nn.ModuleList([nn.Conv3d(2 * c, c, 3, padding=1, bias=False) for _ in range(n_paths)])
        self.last3d = nn.ModuleList([nn.Conv3d(c, 1, 3, padding=1, bias=False) for _ in range(n_paths)])

    @staticmethod
    def _pick(task_arch, name):
        return 0 if task_arch is None else task_arch[name][0]

    def feature(self, x, task_arch=None, path=None):
        x = torch.relu(self.stem2d[self._pick(task_arch, "stem2d")](x))
        return self.last2d[self._pick(task_arch, "last2d")](x)

    def matching(self, cost, task_arch=None, path=None):
        x = torch.relu(self.stem3d[self._pick(task_arch, "stem3d")](cost))
        return self.last3d[self._pick(task_arch, "last3d")](x)

    def search_feature(self, x, selected_ops):
        return self.feature(x, None)

    def search_matching(self, cost, selected_ops, t):
        return self.matching(cost, None)


def arch(i):
    return {"stem2d": [i], "last2d": [i], "stem3d": [i], "last3d": [i]}
