"""`torchdrug.tasks` placeholder: tasks are outside the hot path."""


def _get_metric_name(metric):
    return str(metric)


def _get_criterion_name(criterion):
    return str(criterion)
