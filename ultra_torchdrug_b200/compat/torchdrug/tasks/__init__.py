"""`torchdrug.tasks` stand-in: just enough for the unmodified reference `ultra/task.py` to import and for its
`_strict_negative` / `_calculate_*_mask` methods to run (golden vectors of tests/golden/make_task_golden.py).
Tasks themselves are outside the hot path."""
from torch import nn


def _get_metric_name(metric):
    return str(metric)


def _get_criterion_name(criterion):
    return str(criterion)


class Task(nn.Module):
    pass


class KnowledgeGraphCompletion(Task):
    """Constructor surface of torchdrug's task [ext-recall]: stores its hyper-parameters as attributes."""

    def __init__(self, model, criterion="bce", metric=("mr", "mrr", "hits@1", "hits@3", "hits@10"), num_negative=128, margin=6,
                 adversarial_temperature=0, strict_negative=True, filtered_ranking=True, fact_ratio=None, sample_weight=True):
        super(KnowledgeGraphCompletion, self).__init__()
        self.model = model
        self.criterion = criterion if isinstance(criterion, dict) else {criterion: 1}
        self.metric = metric
        self.num_negative = num_negative
        self.margin = margin
        self.adversarial_temperature = adversarial_temperature
        self.strict_negative = strict_negative
        self.filtered_ranking = filtered_ranking
        self.fact_ratio = fact_ratio
        self.sample_weight = sample_weight
