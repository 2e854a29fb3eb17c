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


def depth_metric_step_ref(depth_loc, depth_gt, depth_mask, depth_pred, eval_types=('raw', 'median')):
    """DepthMetric._after_step (utils/metric_util.py:311-349) restated: -> {type: {metric: [N]}} plus 'scaling'."""
    import torch.nn.functional as F
    num_cams, num_points = depth_gt.shape
    pred = F.grid_sample(depth_pred.unsqueeze(1), depth_loc.unsqueeze(1) * 2 - 1, mode='bilinear', padding_mode='border',
                         align_corners=True).reshape(num_cams, num_points)
    out = {t: {k: [] for k in ('abs_rel', 'sq_rel', 'rmse', 'rmse_log', 'a1', 'a2', 'a3', 'scaling')} for t in eval_types}
    for cam in range(num_cams):
        g, p = depth_gt[cam][depth_mask[cam]], pred[cam][depth_mask[cam]]
        for t in eval_types:
            s = torch.ones(()) if t == 'raw' else torch.median(g) / torch.median(p)
            m = cal_depth_metric_ref(s * p, g)
            for k, v in m.items():
                out[t][k].append(v)
            out[t]['scaling'].append(s)
    return {t: {k: torch.stack([torch.as_tensor(x, dtype=torch.float32) for x in v]) for k, v in d.items()} for t, d in out.items()}, pred
