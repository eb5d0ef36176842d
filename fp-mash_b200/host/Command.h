// Command.h -- option registry and parser with the reference's surface (mash/src/mash/Command.h,
// Command.cpp:55-168,179-238,341-410): same identifiers, defaults, ranges and error texts; numeric
// arguments are stored as float exactly like the reference (Command.h:50), so e.g. -S above 2^24 is
// rounded before it reaches the hash.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace mash {

class Command {
public:
    struct Option {
        enum Type { Boolean, Number, Integer, Size, File, String } type;
        std::string identifier, category, description, argument, argumentDefault;
        float argumentAsNumber = 0, argumentMin = 0, argumentMax = 0;
        bool active = false;
        Option() : type(Boolean) {}
        Option(Type t, std::string id, std::string cat, std::string desc, std::string def = "", float mn = 0, float mx = 0);
        float getArgumentAsNumber() const { return argumentAsNumber; }
        void setArgument(std::string argumentNew);
    };

    Command();
    virtual ~Command() {}
    void addOption(std::string name, Option option);
    const Option& getOption(std::string name) const { return options.at(name); }
    bool hasOption(std::string name) const { return options.count(name) != 0; }
    void print() const;
    int run(int argc, const char** argv);
    virtual int run() const = 0;

    std::string name, summary, description, argumentString;

protected:
    void useOption(std::string name) { addOption(name, optionsAvailable.at(name)); }
    void useSketchOptions();
    std::map<std::string, Option> options, optionsAvailable;
    std::map<std::string, std::string> optionNamesByIdentifier;
    std::vector<std::string> optionOrder;
    std::vector<std::string> arguments;
};

void splitFile(const std::string& file, std::vector<std::string>& lines);

class CommandSketch : public Command { public: CommandSketch(); int run() const; };
class CommandDistance : public Command { public: CommandDistance(); int run() const; };
class CommandPaste : public Command { public: CommandPaste(); int run() const; };
class CommandInfo : public Command { public: CommandInfo(); int run() const; };
class CommandTriangle : public Command { public: CommandTriangle(); int run() const; };
class CommandFingerprint : public Command { public: CommandFingerprint(); int run() const; };

}  // namespace mash
