class DictConfig(dict):
    pass


class OmegaConf:
    @staticmethod
    def create(d):
        return DictConfig(d)
