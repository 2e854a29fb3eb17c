"""Oracle: depth metrics (test infrastructure).  Follows utils/metric_util.py:247-279."""
import torch


def cal_depth_metric_ref(depth_pred, depth_gt):
    depth_pred = torch.clamp(depth_pred, 1e-3, 80)
    thresh = torch.maximum(depth_gt / depth_pred, depth_pred / depth_gt)
    m = lambda t: t.to(torch.float).mean()
    return dict(
        abs_rel=(torch.abs(depth_gt - depth_pred) / depth_gt).mean(),
        sq_rel=(((depth_gt - depth_pred) ** 2) / depth_gt).mean(),
        rmse=((depth_gt - depth_pred) ** 2).mean() ** .5,
        rmse_log=((torch.log(depth_gt) - torch.log(depth_pred)) ** 2).mean() ** .5,
        a1=m(thresh < 1.25), a2=m(thresh < 1.25 ** 2), a3=m(thresh < 1.25 ** 3))
