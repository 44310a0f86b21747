"""lvsr.config of the reference (lvsr/config.py:9-92): YAML with `parent` links, command-line changes and
multi-stage training configurations.  Same functions and semantics; pykwalify validation is skipped when the
schema or pykwalify is unavailable (it only validates)."""
import copy
import logging
import os.path
from collections import OrderedDict

import yaml

logger = logging.getLogger(__name__)


def _load(text_or_file):
    # the reference calls yaml.load(file_) of PyYAML 3: python tags (blocks.bricks.Maxout ...) must resolve
    return yaml.load(text_or_file, Loader=yaml.UnsafeLoader)


def read_config(file_):
    """Reads a configuration from YAML file, resolving parent links (lvsr/config.py:9-22)."""
    config = _load(file_)
    if 'parent' in config:
        with open(os.path.expandvars(config['parent'])) as src:
            changes = dict(config)
            config = read_config(src)
            merge_recursively(config, changes)
    return config


def merge_recursively(config, changes):
    """Merge hierarchy of changes into a configuration (lvsr/config.py:25-31)."""
    for key, value in changes.items():
        if isinstance(value, dict) and isinstance(config.get(key), dict):
            merge_recursively(config[key], value)
        else:
            config[key] = value


def make_config_changes(config, changes):
    """Apply (hierarchical path, new value) pairs (lvsr/config.py:34-49)."""
    for path, value in changes:
        parts = path.split('.')
        assign_to = config
        for part in parts[:-1]:
            assign_to = assign_to[part]
        assign_to[parts[-1]] = _load(value)


class Configuration(dict):
    """lvsr/config.py:52-92: `multi_stage`, `ordered_stages`."""

    def __init__(self, config_path, schema_path, config_changes):
        with open(config_path, 'rt') as src:
            config = read_config(src)
        make_config_changes(config, config_changes)

        self.multi_stage = 'stages' in config
        if self.multi_stage:
            stages = [(k, v) for k, v in config['stages'].items() if v]
            ordered_changes = OrderedDict(sorted(stages, key=lambda kv: kv[1]['number']))
            self.ordered_stages = OrderedDict()
            for name, changes in ordered_changes.items():
                current_config = copy.deepcopy(config)
                del current_config['stages']
                del changes['number']
                merge_recursively(current_config, changes)
                self.ordered_stages[name] = current_config

        if schema_path:
            schema_file = os.path.expandvars(schema_path)
            try:
                from pykwalify.core import Core
                with open(schema_file) as f:
                    schema = yaml.safe_load(f)
                Core(source_data=config, schema_data=schema).validate(raise_exception=True)
                if self.multi_stage:
                    for stage in self.ordered_stages.values():
                        Core(source_data=stage, schema_data=schema).validate(raise_exception=True)
            except (ImportError, IOError, OSError) as e:
                logger.info("configuration not validated (%s)", e)
        super(Configuration, self).__init__(config)
