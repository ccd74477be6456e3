from research.harness.experiment import ExperimentSpec, run_experiment, sweep
from research.harness.methods import METHODS
from research.harness.selection import evaluate_on_test, find_best_hp
from research.harness.servers import FullExchangeServer, PersonalServer
from research.harness.tasks import TASKS, Task

__all__ = ["ExperimentSpec", "FullExchangeServer", "METHODS", "PersonalServer", "TASKS", "Task", "evaluate_on_test",
           "find_best_hp", "run_experiment", "sweep"]
