"""Depth metrics -- drop-in for the reference's ``metrics.py`` (``Result``, ``AverageMeter`` and the confidence
precision/recall meters).  ``Result.evaluate`` (metrics.py:38-66) computes every metric over ``target > 0`` in one
masked pass on the device and reads the ten scalars back with a single transfer (the reference syncs ~12 times)."""
import math

import numpy as np
import torch


class Result(object):
    def __init__(self, threshold=0):
        self.irmse, self.imae = 0, 0
        self.mse, self.rmse, self.mae = 0, 0, 0
        self.absrel, self.lg10 = 0, 0
        self.delta1, self.delta2, self.delta3 = 0, 0, 0
        self.data_time, self.gpu_time = 0, 0
        self.loss0, self.loss1, self.loss2 = 0, 0, 0
        self.threshold = threshold

    def set_to_worst(self):
        self.irmse, self.imae = np.inf, np.inf
        self.mse, self.rmse, self.mae = np.inf, np.inf, np.inf
        self.absrel, self.lg10 = np.inf, np.inf
        self.delta1, self.delta2, self.delta3 = 0, 0, 0
        self.data_time, self.gpu_time = 0, 0
        self.loss0, self.loss1, self.loss2 = np.inf, np.inf, np.inf

    def update(self, irmse, imae, mse, rmse, mae, absrel, lg10, delta1, delta2, delta3, gpu_time, data_time, loss0, loss1, loss2):
        self.irmse, self.imae = irmse, imae
        self.mse, self.rmse, self.mae = mse, rmse, mae
        self.absrel, self.lg10 = absrel, lg10
        self.delta1, self.delta2, self.delta3 = delta1, delta2, delta3
        self.data_time, self.gpu_time = data_time, gpu_time
        self.loss0, self.loss1, self.loss2 = loss0, loss1, loss2

    def evaluate(self, output, target, confidence=None):
        """Every sum in ONE kernel pass, one transfer of the ten scalars (csrc/loss_metrics.cu: adc_depth_metrics).  CUDA tensors
        only: the B200 path has no CPU fallback."""
        from . import ops
        vals = ops.depth_metrics(output.detach().float(), target.detach().float(),
                                 None if confidence is None else confidence.detach().float(), self.threshold).tolist()
        self.mse, self.mae, self.lg10, self.absrel, self.delta1, self.delta2, self.delta3, imse, self.imae = vals[:9]
        self.rmse = math.sqrt(self.mse) if self.mse == self.mse else float('nan')
        self.irmse = math.sqrt(imse) if imse == imse else float('nan')
        self.data_time = 0
        self.gpu_time = 0


class ConfidencePixelwiseAverageMeter(object):
    def __init__(self, num_bins=1000):
        self.num_bins = num_bins
        self.reset()

    def reset(self):
        self.count = np.zeros([self.num_bins], np.uint64)
        self.absrel = np.zeros([self.num_bins], np.uint64)

    def hash_index(self, confidence):
        return np.floor((confidence - 10e-15) * self.num_bins).astype(int)

    def evaluate(self, depth, confidence, target):
        valid_mask = target > 0
        depth, target, confidence = depth[valid_mask], target[valid_mask], confidence[valid_mask]
        indexes = self.hash_index(confidence.cpu().numpy())
        absrel = (((depth - target).abs() / target) * 1000).cpu().numpy()
        np.add.at(self.count, indexes, np.uint64(1))
        np.add.at(self.absrel, indexes, absrel.astype(np.uint64))

    def result(self):
        res = [None] * self.num_bins
        for i in range(self.num_bins):
            if self.count[i] > 0:
                res[i] = self.absrel[i] / self.count[i]
        return res


class ConfidencePixelwiseThrAverageMeter(object):
    def __init__(self, num_bins=500, top=1.0):
        self.num_bins = num_bins
        self.thresholds = np.linspace(0, top, num_bins, endpoint=True)
        self._thr_dev = {}
        self.reset()

    def reset(self):
        self.count = np.zeros([self.num_bins], np.uint64)
        self.absrel = np.zeros([self.num_bins], np.uint64)
        self.recall = np.zeros([self.num_bins], np.uint64)

    def evaluate(self, depth, confidence, target):
        """metrics.py:115-131 loops over the thresholds with a boolean-index pass and two host reads each (1 000 synchronisations
        per call at the default 500 bins).  Here: one histogram kernel over the pixels (adc_conf_pr_histogram), one 12 KB transfer,
        suffix sums on the host.  Threshold comparison in fp32 like the reference's ``confidence > curr_conf``; the masked mean is
        formed from exact integer sums (the reference's is an fp32 reduction), so the truncated ``int(mean)`` can differ by one
        unit when the mean sits within fp32 round-off of an integer."""
        from . import ops
        thr = self._thr_dev.get(confidence.device)
        if thr is None:
            thr = self._thr_dev[confidence.device] = torch.as_tensor(self.thresholds, dtype=torch.float32, device=confidence.device)
        hist = ops.conf_pr_histogram(depth.detach().float(), confidence.detach().float(), target.detach().float(), thr).cpu().numpy()
        cnt_bin, sum_bin, bad_bin = hist[0], hist[1], hist[2]
        total = np.float32(int(cnt_bin.sum()))                             # valid_mask.sum()
        # `valid_mask.sum() / 1000.0` on a CUDA tensor: ATen divides by a host scalar as a multiplication with its fp32 reciprocal
        num_valids = total * (np.float32(1.0) / np.float32(1000.0))
        # threshold j sees the bins k > j
        cnt = np.cumsum(cnt_bin[::-1])[::-1][1:]
        ssum = np.cumsum(sum_bin[::-1])[::-1][1:]
        bad = np.cumsum(bad_bin[::-1])[::-1][1:]
        for i in range(self.num_bins):
            if cnt[i] == 0 or bad[i] > 0:                                  # mean of an empty selection / with a non-finite term: not finite
                continue
            mean = float(ssum[i]) / 65536.0 / float(cnt[i])
            with np.errstate(divide='ignore', invalid='ignore'):
                valids = np.float32(int(cnt[i])) / num_valids
            if math.isfinite(mean) and math.isfinite(valids):
                self.count[i] += 1
                self.absrel[i] += int(mean)
                self.recall[i] += int(valids)

    def result(self):
        res = [(None, None, None)] * self.num_bins
        for i in range(self.num_bins):
            if self.count[i] > 0:
                res[i] = (self.absrel[i] / self.count[i], self.recall[i] / self.count[i], self.thresholds[i])
        return res

    def print(self, filename):
        with open(filename, 'w') as csvfile:
            for absrel, recall, thr in self.result():
                if recall is not None:
                    csvfile.write('{},{},{}\n'.format(thr, recall / 1000.0, absrel / 1000.0))


class AverageMeter(object):
    _FIELDS = ('irmse', 'imae', 'mse', 'rmse', 'mae', 'absrel', 'lg10', 'delta1', 'delta2', 'delta3')

    def __init__(self):
        self.reset()

    def reset(self):
        self.count = 0.0
        for f in self._FIELDS:
            setattr(self, 'sum_' + f, 0)
        self.sum_data_time, self.sum_gpu_time = 0, 0
        self.sum_loss0, self.sum_loss1, self.sum_loss2 = 0, 0, 0

    def update(self, result, gpu_time, data_time, loss, n=1):
        self.count += n
        for f in self._FIELDS:
            setattr(self, 'sum_' + f, getattr(self, 'sum_' + f) + n * getattr(result, f))
        self.sum_data_time += n * data_time
        self.sum_gpu_time += n * gpu_time
        self.sum_loss0 += n * loss[0]
        self.sum_loss1 += n * loss[1]
        self.sum_loss2 += n * loss[2]

    def average(self):
        avg = Result()
        c = self.count
        avg.update(self.sum_irmse / c, self.sum_imae / c, self.sum_mse / c, self.sum_rmse / c, self.sum_mae / c,
                   self.sum_absrel / c, self.sum_lg10 / c, self.sum_delta1 / c, self.sum_delta2 / c, self.sum_delta3 / c,
                   self.sum_gpu_time / c, self.sum_data_time / c, self.sum_loss0 / c, self.sum_loss1 / c, self.sum_loss2 / c)
        return avg
