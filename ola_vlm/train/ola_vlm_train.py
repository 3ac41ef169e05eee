"""Drop-in path of the pieces of ola_vlm/train/ola_vlm_train.py that sit on the training-step path:
the dataset / collator / data module (:774-937), the prompt preprocessing (:350-548) and the
end-of-run save (:228-263)."""
from visper_lm_b200.train.checkpoint import safe_save_model_for_hf_trainer  # noqa: F401
from visper_lm_b200.train.data import (DataArguments, DataCollatorForSupervisedDataset,  # noqa: F401
                                       LazySupervisedDataset, make_supervised_data_module)
from visper_lm_b200.train.prompts import (preprocess_llama_3, preprocess_multimodal,  # noqa: F401,E402
                                          preprocess_phi_3)
from visper_lm_b200.train.entry import ModelArguments, train  # noqa: F401,E402  (ola_vlm_train.py:55-109, 977)
from visper_lm_b200.train.trainer import TrainingArguments  # noqa: F401,E402

if __name__ == "__main__":
    train()
