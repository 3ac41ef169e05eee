from visper_lm_b200.train.trainer import LLaVATrainer, TrainingArguments  # noqa: F401
