def recursively_apply(func, data, *args, test_type=lambda t: True, error_on_other_type=False, **kwargs):
    if isinstance(data, (tuple, list)):
        return type(data)(recursively_apply(func, o, *args, test_type=test_type, **kwargs) for o in data)
    if isinstance(data, dict):
        return type(data)({k: recursively_apply(func, v, *args, test_type=test_type, **kwargs) for k, v in data.items()})
    if test_type(data):
        return func(data, *args, **kwargs)
    return data
