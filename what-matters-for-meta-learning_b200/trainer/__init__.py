"""Drop-in ``trainer.losses`` (LossFunc) for the reference's ``train.py:92`` / ``model_trainer.py:77``."""
