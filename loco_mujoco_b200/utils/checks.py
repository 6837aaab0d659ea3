"""Task-id validation (reference: utils/checks.py:1-76): same acceptance rule, same exception type."""


def check_validity_task_mode_dataset(env_name, task=None, mode=None, dataset_type=None, valid_tasks=None,
                                     valid_modes=None, valid_dataset_types=None, non_combineable=None):
    parts = [p for p, v in (("<task>", task), ("<mode>", mode), ("<dataset_type>", dataset_type)) if v is not None]
    hint = "\n\nThe general structure for calling the environment %s is:\n%s.%s" % (env_name, env_name, ".".join(parts))
    if task is not None:
        hint += "\nValid tasks are %s." % (valid_tasks,)
    if mode is not None:
        hint += "\nValid modes are %s." % (valid_modes,)
    if dataset_type is not None:
        hint += "\nValid dataset types are %s." % (valid_dataset_types,)
    if task is not None and task not in valid_tasks:
        raise ValueError('Task "%s" does not exit in the environment %s. Please, choose from %s. %s'
                         % (task, env_name, valid_tasks, hint))
    if mode is not None and mode not in valid_modes:
        raise ValueError('Mode "%s" does not exit in the environment %s. Please, choose from %s. %s'
                         % (mode, env_name, valid_modes, hint))
    if dataset_type is not None and dataset_type not in valid_dataset_types:
        raise ValueError('Dataset type "%s" does not exit in the environment %s. Please, choose from %s. %s'
                         % (dataset_type, env_name, valid_dataset_types, hint))
    for bad_t, bad_m, bad_dt in (non_combineable or []):
        if (task == bad_t or bad_t is None) and (mode == bad_m or bad_m is None) and \
                (dataset_type == bad_dt or bad_dt is None):
            raise ValueError("Task %r, mode %r and dataset type %r are not combineable for the environment %s. %s"
                             % (task, mode, dataset_type, env_name, hint))
