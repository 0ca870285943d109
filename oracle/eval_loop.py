"""TEST INFRASTRUCTURE ONLY -- restatement of the accumulation and reporting half of the reference's evaluation loop
(core/test.py:141-262): per-sample IoU / F-score per threshold from logits and ground truth (batch size 1, float
tensors, the reference's epsilons and union == 0 convention), per-taxonomy means, sample-weighted overall means and
the two printed tables.  The network half (core/test.py:120-139) is oracle/modules.py.

Pinned by construction: every formula below is the reference's own torch expression on the same tensors; the
printed layout is checked against the reference's print statements character by character in tests/test_evaluate.py.
"""
import io

import numpy as np
import torch


def sample_scores(logits, gt, thresholds):
    """core/test.py:141-164 for one sample -> (iou list, fscore list)"""
    prob = torch.sigmoid(logits)
    ious, fs = [], []
    for th in thresholds:
        vol = torch.ge(prob, th).float()
        inter = torch.sum(vol.mul(gt)).float()
        union = torch.sum(torch.ge(vol.add(gt), 1)).float()
        ious.append(1.0 if union.item() == 0 and inter.item() == 0 else (inter / union).item() if union.item() > 0 else 0.0)
        tp = torch.sum(vol * gt).float()
        fp = torch.sum(vol * (1 - gt)).float()
        fn = torch.sum((1 - vol) * gt).float()
        precision = tp / (tp + fp + 1e-8)
        recall = tp / (tp + fn + 1e-8)
        fs.append((2 * precision * recall / (precision + recall + 1e-8)).item())
    return ious, fs


def sample_losses(merged_logits, refined_logits, gt):
    """core/test.py:133-139: BCEWithLogits * 10 of the merged and of the refined volume"""
    bce = torch.nn.BCEWithLogitsLoss()
    return (bce(merged_logits, gt) * 10).item(), (bce(refined_logits, gt) * 10).item()


def accumulate(taxonomy_ids, all_logits, all_gt, thresholds):
    """core/test.py:166-176,208-220 -> (test_iou, test_fscore, mean_iou, mean_fscore) with the reference's dict layout"""
    test_iou, test_fscore = {}, {}
    for tid, lg, gt in zip(taxonomy_ids, all_logits, all_gt):
        iou, fsc = sample_scores(lg, gt, thresholds)
        test_iou.setdefault(tid, {"n_samples": 0, "iou": []})
        test_iou[tid]["n_samples"] += 1
        test_iou[tid]["iou"].append(iou)
        test_fscore.setdefault(tid, {"n_samples": 0, "fscore": []})
        test_fscore[tid]["n_samples"] += 1
        test_fscore[tid]["fscore"].append(fsc)
    n = len(taxonomy_ids)
    mean_iou, mean_f = [], []
    for tid in test_iou:
        test_iou[tid]["iou"] = np.mean(test_iou[tid]["iou"], axis=0)
        mean_iou.append(test_iou[tid]["iou"] * test_iou[tid]["n_samples"])
    for tid in test_fscore:
        test_fscore[tid]["fscore"] = np.mean(test_fscore[tid]["fscore"], axis=0)
        mean_f.append(test_fscore[tid]["fscore"] * test_fscore[tid]["n_samples"])
    return test_iou, test_fscore, np.sum(mean_iou, axis=0) / n, np.sum(mean_f, axis=0) / n


def tables(test_iou, test_fscore, mean_iou, mean_fscore, taxonomies, thresholds, n_views):
    """the text core/test.py:222-262 prints"""
    out = io.StringIO()

    def p(*a, end="\n"):
        print(*a, end=end, file=out)

    p('============================ TEST RESULTS (IoU) ============================')
    p('Taxonomy', end='\t')
    p('#Sample', end='\t')
    p('Baseline', end='\t')
    for th in thresholds:
        p(f't={th:.2f}', end='\t')
    p()
    for tid in test_iou:
        p(f'{taxonomies[tid]["taxonomy_name"].ljust(8)}', end='\t')
        p(f'{test_iou[tid]["n_samples"]}', end='\t')
        if 'baseline' in taxonomies[tid]:
            key = f"{n_views}-view"
            if key in taxonomies[tid]["baseline"]:
                p(f'{taxonomies[tid]["baseline"][key]:.4f}', end='\t\t')
            else:
                p('N/a', end='\t\t')
        else:
            p('N/a', end='\t\t')
        for ti in test_iou[tid]['iou']:
            p(f'{ti:.4f}', end='\t')
        p()
    p('Overall ', end='\t\t\t\t')
    for mi in mean_iou:
        p(f'{mi:.4f}', end='\t')
    p('\n')
    p('========================== TEST RESULTS (F-score) ==========================')
    p('Taxonomy', end='\t')
    p('#Sample', end='\t')
    p('Baseline', end='\t')
    for th in thresholds:
        p(f't={th:.2f}', end='\t')
    p()
    for tid in test_fscore:
        p(f'{taxonomies[tid]["taxonomy_name"].ljust(8)}', end='\t')
        p(f'{test_fscore[tid]["n_samples"]}', end='\t')
        p('N/a', end='\t\t')
        for sf in test_fscore[tid]['fscore']:
            p(f'{sf:.4f}', end='\t')
        p()
    p('Overall ', end='\t\t\t\t')
    for mf in mean_fscore:
        p(f'{mf:.4f}', end='\t')
    p('\n')
    return out.getvalue()
